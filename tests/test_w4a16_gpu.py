"""GPU parity: W4A16 GEMM on the ZLW4 layout vs the CPU oracle (BASELINE.json config 1 shape and others).

Tolerance (north_star): relative L2 error <= 1e-3 against the fp32 oracle for fp16 outputs."""
import numpy as np
import pytest
import torch

from oracle import gptq, ops as oops
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def _setup(cuda, k, n, sym, seed, row_map=None):
    from zhilight_b200 import ops
    qw, qz, sc, gi = gptq.make_gptq_checkpoint(k, n, 128, sym, seed=seed)
    o_qw, o_qz, o_sc, _ = gptq.to_k_major(qw, qz, sc, gi, 128)
    w = gptq.dequant_k_major_f32(o_qw, o_qz, o_sc, sym)
    packed = ops.w4_pack(_dev(o_qw.view(np.int32), cuda), _dev(o_qz, cuda), _dev(o_sc, cuda), 128, sym, row_map)
    return w, packed, (o_qw, o_qz, o_sc)


@pytest.mark.parametrize("sym", [False, True])
def test_config1_1x4096x4096(lib, cuda, sym):
    """M=1, K=N=4096, g128, seed 0; unpack bit-exact (test_layout) and y rel-err <= 1e-3."""
    from zhilight_b200 import ops
    k = n = 4096
    w, packed, km = _setup(cuda, k, n, sym, 0)
    x = torch.randn(1, k, generator=torch.Generator().manual_seed(0)).half()
    y = ops.w4a16_gemm(x.to(cuda), packed, n, k).float().cpu().numpy()
    ref = gptq.gemm_f32(x.numpy(), w)
    assert rel_l2(y, ref) <= TOL
    # the HF-checkpoint formula of config 1 gives the same weights (zero + 1 with wrap)
    if not sym:
        qw, qz, sc, gi = gptq.make_gptq_checkpoint(k, n, 128, sym, seed=0)
        np.testing.assert_array_equal(w, gptq.hf_dequant_f32(qw, qz, sc, gi, 128))


@pytest.mark.parametrize("m", [1, 2, 3, 7, 8, 9, 16, 17, 31, 32, 33, 70])
def test_all_batch_sizes(lib, cuda, m):
    from zhilight_b200 import ops
    k, n = 1024, 512
    w, packed, _ = _setup(cuda, k, n, False, 5)
    x = torch.randn(m, k, generator=torch.Generator().manual_seed(m)).half()
    y = ops.w4a16_gemm(x.to(cuda), packed, n, k).float().cpu().numpy()
    ref = gptq.gemm_f32(x.numpy(), w)
    assert rel_l2(y, ref) <= TOL
    # row-wise too: no token row may be mixed with another
    for i in range(m):
        assert rel_l2(y[i], ref[i]) <= 2 * TOL


@pytest.mark.parametrize("k,n", [(128, 32), (256, 64), (14336, 4096), (4096, 6144), (4096 + 128, 96)])
def test_shapes_incl_uneven_group_split(lib, cuda, k, n):
    from zhilight_b200 import ops
    w, packed, _ = _setup(cuda, k, n, False, 6)
    x = torch.randn(4, k, generator=torch.Generator().manual_seed(1)).half()
    y = ops.w4a16_gemm(x.to(cuda), packed, n, k).float().cpu().numpy()
    assert rel_l2(y, gptq.gemm_f32(x.numpy(), w)) <= TOL


def test_bias_and_residual_epilogue(lib, cuda):
    from zhilight_b200 import ops
    k, n, m = 512, 256, 5
    w, packed, _ = _setup(cuda, k, n, False, 7)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(m, k, generator=g).half()
    bias = torch.randn(n, generator=g).half()
    res = torch.randn(m, n, generator=g).half()
    y = ops.w4a16_gemm(x.to(cuda), packed, n, k, bias=bias.to(cuda)).float().cpu().numpy()
    ref = gptq.gemm_f32(x.numpy(), w, bias.numpy())
    assert rel_l2(y, ref) <= TOL
    y2 = ops.w4a16_gemm(x.to(cuda), packed, n, k, bias=bias.to(cuda), residual=res.to(cuda),
                        epilogue=ops.EPI_RESIDUAL).float().cpu().numpy()
    ref2 = oops.residual_add(oops._t(ref, "f16"), res.numpy(), "f16")
    assert rel_l2(y2, ref2) <= TOL
    # in-place residual (y aliases residual) as the decode driver uses it
    buf = res.to(cuda).clone()
    ops.w4a16_gemm(x.to(cuda), packed, n, k, bias=bias.to(cuda), residual=buf, epilogue=ops.EPI_RESIDUAL, out=buf)
    np.testing.assert_array_equal(buf.float().cpu().numpy(), y2)


@pytest.mark.parametrize("m", [1, 4, 20])
def test_swiglu_epilogue(lib, cuda, m):
    """silu(x W_gate^T) * (x W_up^T) with gate/up rows interleaved by the packer
    (reference gemm_fuse_gate_in, q_gemm_k_major.cu:529-578; unfused order feedforward.cpp:126-133)."""
    from zhilight_b200 import ops
    k, f = 512, 192
    rm = ops.swiglu_row_map(f, cuda)
    w, packed, _ = _setup(cuda, k, 2 * f, False, 8, rm)
    x = torch.randn(m, k, generator=torch.Generator().manual_seed(3)).half()
    y = ops.w4a16_gemm(x.to(cuda), packed, 2 * f, k, epilogue=ops.EPI_SWIGLU).float().cpu().numpy()
    full = gptq.gemm_f32(x.numpy(), w)
    ref = oops.silu_mul(full[:, :f], full[:, f:], "f16")
    assert y.shape == (m, f)
    assert rel_l2(y, ref) <= 2 * TOL


def test_close_to_reference_kernel_numerics(lib, cuda):
    """The reference GEMV accumulates 8 products in fp16 (q_gemm_k_major.cu:101-108); our fp32 accumulation
    must sit between it and the exact result."""
    from zhilight_b200 import ops
    k, n = 256, 64
    w, packed, (o_qw, o_qz, o_sc) = _setup(cuda, k, n, False, 9)
    x = torch.randn(2, k, generator=torch.Generator().manual_seed(4)).half()
    y = ops.w4a16_gemm(x.to(cuda), packed, n, k).float().cpu().numpy()
    emu = gptq.gemv_ref_numerics(x.numpy(), o_qw, o_qz, o_sc).astype(np.float32)
    exact = gptq.gemm_f32(x.numpy(), w)
    assert rel_l2(y, exact) <= rel_l2(emu, exact) + 5e-4
    assert rel_l2(y, emu) <= 3e-3


def test_pdl_launch_gives_same_result(lib, cuda):
    from zhilight_b200 import ops
    k, n = 1024, 256
    w, packed, _ = _setup(cuda, k, n, False, 10)
    x = torch.randn(3, k, generator=torch.Generator().manual_seed(5)).half().to(cuda)
    a = ops.w4a16_gemm(x, packed, n, k)
    b = ops.w4a16_gemm(x, packed, n, k, pdl=True)
    c = ops.w4a16_gemm(x, packed, n, k, pdl=True)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(a, c)


def test_rejects_bf16_like_reference(lib, cuda):
    from zhilight_b200 import ops, ZLError
    w, packed, _ = _setup(cuda, 128, 32, False, 11)
    with pytest.raises(ZLError):
        ops.w4a16_gemm(torch.zeros(1, 128, dtype=torch.bfloat16, device=cuda), packed, 32, 128)


@pytest.mark.parametrize("m", [1, 5, 12, 40])
def test_fused_rmsnorm_prologue(lib, cuda, m):
    """y = W . rmsnorm(x) with the norm folded into the GEMM (zl_w4a16_gemm_fused, ln_weight != NULL)."""
    from zhilight_b200 import ops
    k, n = 1024, 320
    w, packed, _ = _setup(cuda, k, n, False, 12)
    g = torch.Generator().manual_seed(6)
    x = (torch.randn(m, k, generator=g) * 3.0).half()
    lw = (1 + 0.1 * torch.randn(k, generator=g)).half()
    y = ops.w4a16_gemm_fused(x.to(cuda), packed, n, k, ln_weight=lw.to(cuda), eps=1e-5).float().cpu().numpy()
    xn = oops.rmsnorm(x.numpy(), lw.numpy(), 1e-5, 1.0, "f16")
    ref = gptq.gemm_f32(xn, w)
    assert rel_l2(y, ref) <= TOL
    # and it matches the two-kernel path to the same tolerance
    y2 = ops.w4a16_gemm(ops.rmsnorm(x.to(cuda), lw.to(cuda), 1e-5), packed, n, k).float().cpu().numpy()
    assert rel_l2(y, y2) <= TOL


def test_fused_rmsnorm_with_swiglu_and_many_tiles(lib, cuda):
    """N large enough that persistent CTAs walk several super-tiles (grid = 2 per SM)."""
    from zhilight_b200 import ops
    k, f = 512, 5120                       # 2f/32 = 320 super-tiles > 296 CTAs
    rm = ops.swiglu_row_map(f, cuda)
    w, packed, _ = _setup(cuda, k, 2 * f, False, 13, rm)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, k, generator=g).half()
    lw = (1 + 0.1 * torch.randn(k, generator=g)).half()
    y = ops.w4a16_gemm_fused(x.to(cuda), packed, 2 * f, k, epilogue=ops.EPI_SWIGLU, ln_weight=lw.to(cuda),
                             eps=1e-5).float().cpu().numpy()
    xn = oops.rmsnorm(x.numpy(), lw.numpy(), 1e-5, 1.0, "f16")
    full = gptq.gemm_f32(xn, w)
    ref = oops.silu_mul(full[:, :f], full[:, f:], "f16")
    assert rel_l2(y, ref) <= 2 * TOL


@pytest.mark.parametrize("d,hq,hkv", [(128, 4, 2), (64, 4, 1)])
def test_fused_qkv_rope_epilogue(lib, cuda, d, hq, hkv):
    """qkv GEMM + split + RoPE + KV append in one kernel == GEMM, then rope_qk_cache, then copy_to_rag_buffer2."""
    from zhilight_b200 import ops
    k = 512
    n = (hq + 2 * hkv) * d
    rm = ops.qkv_rope_row_map(hq + 2 * hkv, d, cuda)
    w, packed, _ = _setup(cuda, k, n, False, 14, rm)
    w_plain, packed_plain, _ = _setup(cuda, k, n, False, 14)
    g = torch.Generator().manual_seed(8)
    t = 5
    x = torch.randn(t, k, generator=g).half().to(cuda)
    bias = torch.randn(n, generator=g).half().to(cuda)
    pos = torch.tensor([0, 3, 7, 2, 9], dtype=torch.int32, device=cuda)
    cos, sin = ops.rope_cos_sin(pos, d, 10000.0)
    tb = torch.tensor([0, 1, 2, 1, 0], dtype=torch.int32, device=cuda)
    pl = torch.tensor([0, 3, 7, 2, -1], dtype=torch.int32, device=cuda)
    cap = 12
    kb = [torch.zeros(cap, hkv, d, dtype=torch.float16, device=cuda) for _ in range(3)]
    vb = [torch.zeros(cap, hkv, d, dtype=torch.float16, device=cuda) for _ in range(3)]
    q = ops.w4a16_gemm_fused(x, packed, n, k, bias=ops.gather_rows_16(bias, rm), epilogue=ops.EPI_QKV_ROPE,
                             rope=dict(cos=cos, sin=sin, token_batch=tb, placement=pl, k_bufs=kb, v_bufs=vb,
                                       num_heads=hq, num_kv_heads=hkv, dim_head=d))
    # unfused path on the same weights (natural row order)
    qkv = ops.w4a16_gemm(x, packed_plain, n, k, bias=bias)
    kb2 = [torch.zeros_like(b) for b in kb]
    vb2 = [torch.zeros_like(b) for b in vb]
    q2 = ops.qkv_rope_append(cos, sin, qkv, tb, pl, kb2, vb2, hq, hkv, d)
    torch.testing.assert_close(q, q2, rtol=2 ** -9, atol=2e-3)
    for a, b in zip(kb + vb, kb2 + vb2):
        torch.testing.assert_close(a, b, rtol=2 ** -9, atol=2e-3)
    assert kb[0][0].abs().sum() > 0 and vb[2][7].abs().sum() > 0
