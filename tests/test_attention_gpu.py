"""GPU parity: batch decode attention over per-task ragged KV buffers vs the oracle
(reference semantics attention_kernel.cu:434-489, 674-725, 730-923)."""
import numpy as np
import pytest
import torch

from oracle import ops as oops
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _case(cuda, dtype, name, hq, hkv, d, lens, len_q=1, masked=True, seed=0, bshd=True):
    from zhilight_b200 import ops
    g = torch.Generator().manual_seed(seed)
    b = len(lens)
    q = torch.randn(b, len_q, hq, d, generator=g).to(dtype)
    ks, vs, masks = [], [], []
    for lb in lens:
        shape = (lb, hkv, d) if bshd else (hkv, lb, d)
        ks.append(torch.randn(*shape, generator=g).to(dtype))
        vs.append(torch.randn(*shape, generator=g).to(dtype))
        m = torch.ones(len_q, lb, dtype=torch.int8)
        if masked:
            for qi in range(len_q):
                valid = max(1, lb - 2 - 3 * qi)           # the reference's len_buf is padded past the tokens
                m[qi, valid:] = 0
                m[qi, torch.randperm(valid, generator=g)[: valid // 7]] = 0   # beam-tree style holes
        masks.append(m)
    scale = 1.0 / np.sqrt(d)
    mask_flat = torch.cat([m.reshape(-1) for m in masks]).to(cuda)
    out = ops.decode_attention(q.to(cuda), torch.tensor(lens, dtype=torch.int32, device=cuda),
                               [k.to(cuda) for k in ks], [v.to(cuda) for v in vs], mask_flat if masked else None,
                               scale, max(lens), hkv, bshd=bshd)
    kn = [k.float().numpy() if bshd else k.float().permute(1, 0, 2).contiguous().numpy() for k in ks]
    vn = [v.float().numpy() if bshd else v.float().permute(1, 0, 2).contiguous().numpy() for v in vs]
    ref = oops.decode_attention(q.float().numpy(), kn, vn, lens, [m.numpy() for m in masks], scale, hq // hkv, name)
    return out.float().cpu().numpy(), ref


@pytest.mark.parametrize("dtype,name,tol", [(torch.float16, "f16", 1e-3), (torch.bfloat16, "bf16", 6e-3)])
@pytest.mark.parametrize("hq,hkv,d", [(32, 8, 128), (32, 8, 64), (8, 1, 128), (16, 16, 64), (64, 8, 128), (32, 2, 128)])
def test_gqa_shapes(lib, cuda, dtype, name, tol, hq, hkv, d):
    out, ref = _case(cuda, dtype, name, hq, hkv, d, [130, 64, 17, 1])
    assert rel_l2(out, ref) <= tol
    # per (task, head) as well: a wrong head/task mapping cannot hide in the global norm
    for b in range(out.shape[0]):
        for h in range(hq):
            assert rel_l2(out[b, 0, h], ref[b, 0, h]) <= 4 * tol, (b, h)


@pytest.mark.parametrize("lens", [[4096, 123], [2048] * 3, [1500, 700, 3000, 64, 65], [1], [16], [15, 33]])
def test_split_kv_and_ragged_lengths(lib, cuda, lens):
    """src/nn/tests/test_attention_rag_buffer.cpp uses lens {4096, 123}, 4 kv-heads x m_query 4."""
    out, ref = _case(cuda, torch.float16, "f16", 16, 4, 128, lens, seed=1)
    assert rel_l2(out, ref) <= 1e-3


def test_no_mask_and_beam_queries(lib, cuda):
    out, ref = _case(cuda, torch.float16, "f16", 8, 2, 128, [300, 90], len_q=3, masked=True, seed=2)
    assert rel_l2(out, ref) <= 1e-3
    out, ref = _case(cuda, torch.float16, "f16", 8, 2, 128, [300, 90], len_q=1, masked=False, seed=3)
    assert rel_l2(out, ref) <= 1e-3


def test_hsd_layout(lib, cuda):
    out, ref = _case(cuda, torch.float16, "f16", 8, 2, 64, [200, 40], seed=4, bshd=False)
    assert rel_l2(out, ref) <= 1e-3


def test_fully_masked_row_is_zero_not_nan(lib, cuda):
    from zhilight_b200 import ops
    q = torch.randn(1, 1, 4, 64, device=cuda).half()
    k = [torch.randn(32, 1, 64, device=cuda).half()]
    v = [torch.randn(32, 1, 64, device=cuda).half()]
    mask = torch.zeros(32, dtype=torch.int8, device=cuda)
    out = ops.decode_attention(q, torch.tensor([32], dtype=torch.int32, device=cuda), k, v, mask, 0.125, 32, 1)
    assert torch.isfinite(out).all() and out.abs().max() < 1e-6     # sum seeded with 1e-20 (attention_kernel.cu:466)
