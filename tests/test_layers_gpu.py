"""`zhilight_b200.layers.Linear` used the way the reference's tests use `zhilight.internals_.layers.Linear`
(tests/test_linear.py:48-86: same constructor arguments, load_state_dict / named_parameters / forward, same
tolerances fp16 rtol 1e-3 atol 3e-3, bf16 rtol 1e-2), extended over the quantised QuantTypes of SURVEY 8-a1."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import gptq, ops as oracle
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("SIZE", [(64, 32), (4096, 1024)])
@pytest.mark.parametrize("BATCH", [2, 4])
@pytest.mark.parametrize("SEQLEN", [4, 8])
@pytest.mark.parametrize("ACTIVATION", ["", "silu", "gelu"])
@pytest.mark.parametrize("DTYPE", [torch.half, torch.bfloat16])
def test_linear(lib, cuda, SIZE, BATCH, SEQLEN, ACTIVATION, DTYPE):
    from zhilight_b200 import layers
    rtol, atol = (1e-2, 3e-3) if DTYPE == torch.bfloat16 else (1e-3, 3e-3)
    ff = layers.Linear(SIZE[0], SIZE[1], ACTIVATION, False, "bfloat" if DTYPE == torch.bfloat16 else "half")
    g = torch.Generator().manual_seed(0)
    weight_pt = (torch.randn(SIZE[1], SIZE[0], generator=g) / SIZE[0] ** 0.5).to(DTYPE).to(cuda)
    inp = torch.randn([BATCH, SEQLEN, SIZE[0]], generator=g).to(DTYPE).to(cuda)
    ff.load_state_dict({"weight": weight_pt})
    assert torch.equal(ff.named_parameters()["weight"], weight_pt)
    out = ff.forward(inp)
    out_pt = F.linear(inp.float(), weight_pt.float())
    if ACTIVATION:
        out_pt = (F.silu if ACTIVATION == "silu" else F.gelu)(out_pt)
    assert out.shape == (BATCH, SEQLEN, SIZE[1])
    assert torch.allclose(out.float(), out_pt.to(DTYPE).float(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("quant,sym", [(5, False), (5, True), (8, True), (6, False)])
def test_linear_w4(lib, cuda, quant, sym):
    from zhilight_b200 import layers
    k, n = 4096, 1024
    if quant == 6:
        r = np.random.default_rng(1)
        qw = r.integers(0, 2 ** 32, size=(k, n // 8), dtype=np.uint64).astype(np.uint32).view(np.int32)
        qz = r.integers(0, 2 ** 32, size=(k // 128, n // 8), dtype=np.uint64).astype(np.uint32).view(np.int32)
        sc = (0.002 + 0.01 * r.random((k // 128, n))).astype(np.float16)
        sd = {"qweight": qw, "qzeros": qz, "scales": sc}
        km = gptq.to_k_major(qw, qz, sc, None, 128, is_awq=True)
    else:
        qw, qz, sc, gi = gptq.make_gptq_checkpoint(k, n, 128, sym, seed=3)
        sd = {"qweight": qw, "qzeros": qz, "scales": sc, "g_idx": gi}
        km = gptq.to_k_major(qw, qz, sc, gi, 128)
    ff = layers.Linear(k, n, "", quant, "half", sym=sym)
    ff.load_state_dict({kk: torch.from_numpy(np.ascontiguousarray(v)) for kk, v in sd.items()})
    x = torch.randn(3, 11, k, generator=torch.Generator().manual_seed(2)).half()
    y = ff.forward(x.to(cuda)).float().cpu().numpy().reshape(33, n)
    exact = gptq.gemm_f32(x.reshape(33, k).numpy(), gptq.dequant_k_major_f32(km[0], km[1], km[2], sym))
    assert rel_l2(y, exact) <= 1e-3


def test_linear_auto_int8_and_fp8(lib, cuda):
    from zhilight_b200 import layers
    k, n = 1024, 768
    g = torch.Generator().manual_seed(7)
    w = (torch.randn(n, k, generator=g) * 0.05).half()
    x = torch.randn(5, k, generator=g).half()
    ff = layers.Linear(k, n, "", 2, "half")
    ff.load_state_dict({"weight": w})
    wq, ws = oracle.int8_quant_per_token(w.float().numpy())
    np.testing.assert_array_equal(ff.named_parameters()["weight"].cpu().numpy(), wq)
    ref = oracle.int8_linear(x.float().numpy(), wq, oracle._t(ws, "f16"), "f16")
    np.testing.assert_array_equal(ff.forward(x.to(cuda)).float().cpu().numpy(), ref)
    # the int8 Linear stays close to the fp Linear it replaces (SmoothQuant-level error)
    assert rel_l2(ref, (x.float() @ w.float().T).numpy()) < 2e-2

    w8 = (torch.randn(n, k, generator=g) * 0.5).to(torch.float8_e4m3fn)
    f8 = layers.Linear(k, n, "", 7, "half")
    f8.load_state_dict({"weight": w8, "weight_scale": torch.tensor([0.02])})
    ref8 = oracle.fp8_linear(x.float().numpy(), w8.float().numpy(), np.float32(0.02), "f16")
    assert rel_l2(f8.forward(x.to(cuda)).float().cpu().numpy(), ref8) < 1e-3


def test_linear_errors(lib, cuda):
    from zhilight_b200 import _lib, layers
    with pytest.raises(_lib.ZLError):
        layers.Linear(64, 64, "", 3, "half")                      # quant value error
    with pytest.raises(_lib.ZLError):
        layers.Linear(4096, 4096, "", 5, "bfloat")                # "A must be half"
    ff = layers.Linear(64, 32, "", 0, "half")
    ff.load_state_dict({"weight": torch.zeros(32, 64).half()})
    with pytest.raises(_lib.ZLError):
        ff.forward(torch.zeros(4, 32, dtype=torch.half, device=cuda))          # Input size mismatch
    with pytest.raises(_lib.ZLError):
        ff.forward(torch.zeros(4, 64, dtype=torch.bfloat16, device=cuda))      # dtype mismatch
