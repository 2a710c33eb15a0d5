"""GPU parity: load-time integer transforms and the ZLW4 packer are bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from oracle import gptq
from tests.helpers import w4_pack_numpy

pytestmark = pytest.mark.gpu


def _dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


@pytest.mark.parametrize("k,n", [(128, 32), (256, 96), (4096, 4096)])
def test_gptq_pipeline_bit_exact(lib, cuda, k, n):
    from zhilight_b200 import ops
    qw, qz, sc, gi = gptq.make_gptq_checkpoint(k, n, 128, False, seed=k + n)
    o_qw, o_qz, o_sc, _ = gptq.to_k_major(qw, qz, sc, gi, 128)
    # individual stages
    t_qw = _dev(qw, cuda)
    np.testing.assert_array_equal(ops.gptq_shuffle(t_qw.clone()).cpu().numpy().view(np.uint32), gptq.gptq_shuffle(qw))
    np.testing.assert_array_equal(ops.gptq_increase_zero(_dev(qz, cuda)).cpu().numpy().view(np.uint32),
                                  gptq.increase_zero(qz))
    np.testing.assert_array_equal(ops.gptq_subtract8(_dev(qz, cuda)).cpu().numpy().view(np.uint32), gptq.subtract8(qz))
    np.testing.assert_array_equal(ops.q4_to_q8(_dev(qz, cuda)).cpu().numpy(), gptq.q4_to_q8(qz))
    # whole pipeline
    d_qw, d_qz, d_sc = ops.gptq_to_k_major(t_qw, _dev(qz, cuda), _dev(sc, cuda))
    np.testing.assert_array_equal(d_qw.cpu().numpy().view(np.uint32), o_qw)
    np.testing.assert_array_equal(d_qz.cpu().numpy(), o_qz)
    np.testing.assert_array_equal(d_sc.cpu().numpy(), o_sc)
    # dequant-to-fp16 (KERNEL_dequant semantics) bit-exact
    w16 = ops.gptq_dequant_k_major(d_qw, d_qz, d_sc).cpu().numpy()
    np.testing.assert_array_equal(w16, gptq.dequant_k_major_f16(o_qw, o_qz, o_sc))


def test_act_order_shuffle_bit_exact(lib, cuda):
    from zhilight_b200 import ops
    k, n, g = 512, 64, 128
    qw, qz, sc, _ = gptq.make_gptq_checkpoint(k, n, g, False, seed=11)
    g_idx = np.random.default_rng(0).permutation(np.arange(k) // g).astype(np.int32)
    perm = gptq.argsort_g_idx(g_idx, g)
    out = ops.gptq_shuffle(_dev(qw, cuda), _dev(perm, cuda)).cpu().numpy().view(np.uint32)
    np.testing.assert_array_equal(out, gptq.gptq_shuffle(qw, perm))


def test_awq_pipeline_bit_exact(lib, cuda):
    from zhilight_b200 import ops
    k, n, g = 256, 128, 128
    rng = np.random.default_rng(1)
    qw = rng.integers(0, 2 ** 32, size=(k, n // 8), dtype=np.uint64).astype(np.uint32).view(np.int32)
    qz = rng.integers(0, 2 ** 32, size=(k // g, n // 8), dtype=np.uint64).astype(np.uint32).view(np.int32)
    sc = (0.01 * rng.random((k // g, n))).astype(np.float16)
    for exl in (True, False):
        np.testing.assert_array_equal(ops.awq_shuffle(_dev(qw, cuda), exl).cpu().numpy().view(np.uint32),
                                      gptq.shuffle_awq(qw, exl))
    np.testing.assert_array_equal(ops.awq_un_shuffle(_dev(qz, cuda)).cpu().numpy().view(np.uint32), gptq.un_shuffle(qz))
    o_qw, o_qz, o_sc, _ = gptq.to_k_major(qw, qz, sc, None, g, is_awq=True)
    d_qw, d_qz, d_sc = ops.gptq_to_k_major(_dev(qw, cuda), _dev(qz, cuda), _dev(sc, cuda), is_awq=True)
    np.testing.assert_array_equal(d_qw.cpu().numpy().view(np.uint32), o_qw)
    np.testing.assert_array_equal(d_qz.cpu().numpy(), o_qz)
    np.testing.assert_array_equal(d_sc.cpu().numpy(), o_sc)


@pytest.mark.parametrize("rows,cols,dtype", [(5, 7, np.uint8), (33, 65, np.uint16), (128, 96, np.uint32), (1, 40, np.uint32)])
def test_transpose(lib, cuda, rows, cols, dtype):
    from zhilight_b200 import ops
    a = np.random.default_rng(0).integers(0, 255, size=(rows, cols)).astype(dtype)
    t = torch.from_numpy(a.view({1: np.uint8, 2: np.int16, 4: np.int32}[a.itemsize])).to(cuda)
    np.testing.assert_array_equal(ops.transpose_2d(t).cpu().numpy().view(dtype), a.T)


@pytest.mark.parametrize("k,n,sym", [(128, 32, False), (512, 64, True), (4096, 256, False)])
def test_w4_pack_bit_exact_and_roundtrip(lib, cuda, k, n, sym):
    from zhilight_b200 import ops
    qw, qz, sc, gi = gptq.make_gptq_checkpoint(k, n, 128, sym, seed=3)
    o_qw, o_qz, o_sc, _ = gptq.to_k_major(qw, qz, sc, gi, 128)
    packed = ops.w4_pack(_dev(o_qw.view(np.int32), cuda), _dev(o_qz, cuda), _dev(o_sc, cuda), 128, sym)
    np.testing.assert_array_equal(packed.cpu().numpy(), w4_pack_numpy(o_qw, o_qz, o_sc, sym))
    r_qw, r_qz, r_sc = ops.w4_unpack(packed, n, k)
    np.testing.assert_array_equal(r_qw.cpu().numpy().view(np.uint32), o_qw)
    np.testing.assert_array_equal(r_qz.cpu().numpy(), np.full_like(o_qz, 8) if sym else o_qz)
    np.testing.assert_array_equal(r_sc.cpu().numpy(), o_sc)


def test_w4_pack_row_map(lib, cuda):
    from zhilight_b200 import ops
    k, f = 256, 64
    qw, qz, sc, gi = gptq.make_gptq_checkpoint(k, 2 * f, 128, False, seed=4)
    o_qw, o_qz, o_sc, _ = gptq.to_k_major(qw, qz, sc, gi, 128)
    rm = ops.swiglu_row_map(f, cuda)
    packed = ops.w4_pack(_dev(o_qw.view(np.int32), cuda), _dev(o_qz, cuda), _dev(o_sc, cuda), 128, False, rm)
    np.testing.assert_array_equal(packed.cpu().numpy(), w4_pack_numpy(o_qw, o_qz, o_sc, False, rm.cpu().numpy()))
