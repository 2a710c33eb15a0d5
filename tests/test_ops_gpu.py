"""GPU parity: skinny dense GEMM, RMSNorm (+residual), element add, gate-mul, RoPE, KV append, argmax."""
import numpy as np
import pytest
import torch

from oracle import ops as oops
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _tn(t):
    return t.float().cpu().numpy()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("m,n,k", [(1, 512, 256), (5, 1000, 512), (17, 96, 2048), (40, 130, 64)])
def test_dense_gemm_skinny(lib, cuda, dtype, m, n, k):
    from zhilight_b200 import ops
    g = torch.Generator().manual_seed(m * n)
    x = (torch.randn(m, k, generator=g) * 0.5).to(dtype)
    w = (torch.randn(n, k, generator=g) * 0.05).to(dtype)
    b = torch.randn(n, generator=g).to(dtype)
    ref = x.float() @ w.float().T + b.float()
    y32 = ops.dense_gemm_skinny(x.to(cuda), w.to(cuda), b.to(cuda), out_dtype=torch.float32)
    assert rel_l2(_tn(y32), ref.numpy()) <= 1e-5
    y = ops.dense_gemm_skinny(x.to(cuda), w.to(cuda), b.to(cuda))
    tol = 1e-3 if dtype == torch.float16 else 8e-3      # reference tests/test_linear.py:48-86 rtol 1e-3 / 1e-2
    assert rel_l2(_tn(y), ref.numpy()) <= tol


@pytest.mark.parametrize("dtype,name", [(torch.float16, "f16"), (torch.bfloat16, "bf16")])
@pytest.mark.parametrize("t,d", [(1, 4096), (3, 2048), (32, 256), (2, 8192)])
def test_rmsnorm_and_add(lib, cuda, dtype, name, t, d):
    from zhilight_b200 import ops
    g = torch.Generator().manual_seed(d + t)
    a = torch.randn(t, d, generator=g).to(dtype)
    b = torch.randn(t, d, generator=g).to(dtype)
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype)
    y = ops.rmsnorm(a.to(cuda), w.to(cuda), 1e-5)
    ref = oops.rmsnorm(_tn(a), _tn(w), 1e-5, 1.0, name)
    ulp = 2 ** -10 if name == "f16" else 2 ** -7
    np.testing.assert_allclose(_tn(y), ref, rtol=ulp, atol=1e-6)
    # mode 0: reference single-stream order (add in T, then norm)
    s0, y0 = ops.add_rmsnorm(a.to(cuda), b.to(cuda), w.to(cuda), 1e-5, mode=0)
    h, r0 = oops.add_then_rmsnorm(_tn(a), _tn(b), _tn(w), 1e-5, name)
    np.testing.assert_array_equal(_tn(s0), h)
    np.testing.assert_allclose(_tn(y0), r0, rtol=ulp, atol=1e-6)
    # mode 1: LayerNorm::fuse_add (normalises the unrounded sum)
    s1, y1 = ops.add_rmsnorm(a.to(cuda), b.to(cuda), w.to(cuda), 1e-5, mode=1)
    h1, r1 = oops.add_rmsnorm_fused(_tn(a), _tn(b), _tn(w), 1e-5, 1.0, name)
    np.testing.assert_array_equal(_tn(s1), h1)
    np.testing.assert_allclose(_tn(y1), r1, rtol=ulp, atol=1e-6)
    c = ops.element_add_scale(a.to(cuda), b.to(cuda), 1.0)
    np.testing.assert_array_equal(_tn(c), oops.residual_add(_tn(a), _tn(b), name))


def test_gate_mul(lib, cuda):
    from zhilight_b200 import ops
    g = torch.Generator().manual_seed(0)
    gu = torch.randn(6, 2 * 384, generator=g).half().to(cuda)
    out = ops.gate_mul(gu[:, :384], gu[:, 384:], "silu")
    ref = oops.silu_mul(_tn(gu[:, :384]), _tn(gu[:, 384:]), "f16")
    np.testing.assert_allclose(_tn(out), ref, rtol=2 ** -10, atol=1e-6)


@pytest.mark.parametrize("llama3", [None, dict(factor=8.0, low=1.0, high=4.0, orig=8192.0),
                                    dict(factor=32.0, low=1.0, high=4.0, orig=8192.0)])
@pytest.mark.parametrize("d", [64, 128])
def test_rope_tables(lib, cuda, llama3, d):
    from zhilight_b200 import ops
    pos = torch.tensor([0, 1, 17, 511, 4095, 100000], dtype=torch.int32)
    cos, sin = ops.rope_cos_sin(pos.to(cuda), d, 500000.0, llama3)
    rc, rs = oops.rope_cos_sin(pos.numpy(), d, 500000.0, llama3)
    # fp32 powf/cosf vs float64 oracle: the phase m*inv_freq carries ~1e-7 relative error, i.e. up to
    # 1e-7 * 1e5 rad at the largest position
    np.testing.assert_allclose(_tn(cos), rc, atol=3e-2)
    np.testing.assert_allclose(_tn(cos)[:4], rc[:4], atol=2e-4)
    np.testing.assert_allclose(_tn(sin)[:4], rs[:4], atol=2e-4)


@pytest.mark.parametrize("dtype,name", [(torch.float16, "f16"), (torch.bfloat16, "bf16")])
def test_rope_qk_cache_and_fused_append(lib, cuda, dtype, name):
    from zhilight_b200 import ops
    hq, hkv, d, t = 8, 2, 64, 5
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(t, (hq + 2 * hkv) * d, generator=g).to(dtype)
    pos = torch.tensor([3, 0, 9, 1, 2], dtype=torch.int32)
    cos, sin = ops.rope_cos_sin(pos.to(cuda), d, 10000.0)
    q, k, v = ops.rope_qk_cache(cos, sin, qkv.to(cuda), hq, hkv, d)
    rq, rk, rv = oops.split_qkv_rope(_tn(qkv), _tn(cos), _tn(sin), hq, hkv, d, True, name)
    ulp = 2 ** -9 if name == "f16" else 2 ** -6
    np.testing.assert_allclose(_tn(q), rq, rtol=ulp, atol=1e-3)
    np.testing.assert_allclose(_tn(k), rk, rtol=ulp, atol=1e-3)
    np.testing.assert_array_equal(_tn(v), rv)
    # fused variant: same q, K/V land in the per-task buffers at `placement`
    cap = 16
    kb = [torch.zeros(cap, hkv, d, dtype=dtype, device=cuda) for _ in range(3)]
    vb = [torch.zeros(cap, hkv, d, dtype=dtype, device=cuda) for _ in range(3)]
    tb = torch.tensor([0, 1, 2, 1, 0], dtype=torch.int32, device=cuda)
    pl = torch.tensor([3, 0, 9, 1, -1], dtype=torch.int32, device=cuda)
    q2 = ops.qkv_rope_append(cos, sin, qkv.to(cuda), tb, pl, kb, vb, hq, hkv, d)
    assert torch.equal(q2, q)
    kk, vv = k.view(t, hkv, d), v.view(t, hkv, d)
    assert torch.equal(kb[0][3], kk[0]) and torch.equal(kb[1][0], kk[1]) and torch.equal(kb[2][9], kk[2])
    assert torch.equal(vb[1][1], vv[3])
    assert kb[0].abs().sum() == kk[0].abs().sum()          # token 4 (placement -1) was skipped
    # reference-shaped copy_to_rag_buffer2 (B, len_q) placement
    kb2 = [torch.zeros(cap, hkv, d, dtype=dtype, device=cuda) for _ in range(t)]
    vb2 = [torch.zeros(cap, hkv, d, dtype=dtype, device=cuda) for _ in range(t)]
    place = torch.tensor([[3], [0], [9], [1], [2]], dtype=torch.int32, device=cuda)
    lens = torch.full((t,), cap, dtype=torch.int32, device=cuda)
    ops.copy_to_rag_buffer2(place, lens, kk.view(t, 1, hkv, d).contiguous(), vv.view(t, 1, hkv, d).contiguous(), kb2, vb2)
    for i, p in enumerate([3, 0, 9, 1, 2]):
        assert torch.equal(kb2[i][p], kk[i]) and torch.equal(vb2[i][p], vv[i])


def test_embedding_and_argmax(lib, cuda):
    from zhilight_b200 import ops
    g = torch.Generator().manual_seed(2)
    table = torch.randn(300, 64, generator=g).half().to(cuda)
    ids = torch.tensor([5, 299, 0, 17], dtype=torch.int32, device=cuda)
    assert torch.equal(ops.embedding(ids, table), table[ids.long()])
    logits = torch.randn(7, 128256, generator=g).to(cuda)
    logits[3, 77] = logits[3, 99999] = 50.0                # tie -> lowest index
    out = ops.argmax(logits)
    ref = logits.argmax(dim=1)
    assert out[3].item() == 77
    for i in (0, 1, 2, 4, 5, 6):
        assert out[i].item() == ref[i].item()
