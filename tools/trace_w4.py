#!/usr/bin/env python
"""Timeline of the integer W4A16 kernel from per-CTA %globaltimer samples (debug facility zl_w4_set_trace)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zhilight_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


def run(n, k, m, epi, label):
    nbytes = (n // 32) * (k // 128) * 2128
    packs = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=dev) for _ in range(6)]
    for p in packs:
        v = p.view(-1, 2128)
        v[:, 2048:2112] = 0
        v[:, 2049:2112:2] = 0x1c
    x = torch.randn(m, k, device=dev).half()
    res = torch.zeros(m, n, device=dev).half()
    out = torch.empty(m, n // 2 if epi == 1 else n, device=dev).half()
    trace = torch.zeros(1024 * 16, dtype=torch.int64, device=dev)
    for i in range(5):
        ops.w4a16_gemm_fused(x, packs[i], n, k, residual=res if epi == 2 else None, epilogue=epi, out=out, variant=1)
    torch.cuda.synchronize()
    lib.zl_w4_set_trace(ctypes.c_void_p(trace.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.w4a16_gemm_fused(x, packs[5], n, k, residual=res if epi == 2 else None, epilogue=epi, out=out, variant=1)
    e1.record()
    torch.cuda.synchronize()
    lib.zl_w4_set_trace(None)
    t = trace.cpu().numpy().reshape(1024, 16)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t - t0).astype(np.float64) / 1e3
    rel[t == 0] = np.nan
    print("== %s N=%d K=%d M=%d: %d CTAs, event time %.2f us, ideal %.2f us" % (label, n, k, m, len(t), e0.elapsed_time(e1) * 1e3,
                                                                               nbytes / 6.5677e6))
    names = ["entry", "ring issued", "pdl_wait done", "staging done"] + ["tile %d done" % i for i in range(12)]
    for c in range(16):
        col = rel[:, c]
        if np.all(np.isnan(col)):
            break
        print("  %-14s min %6.2f  median %6.2f  max %6.2f us  (n=%d)" % (names[c], np.nanmin(col), np.nanmedian(col), np.nanmax(col),
                                                                       np.sum(~np.isnan(col))))
    print("  kernel span (first entry -> last stamp): %.2f us" % np.nanmax(rel))


run(28672, 4096, 1, 1, "gate_up")
run(4096, 14336, 1, 2, "down")
run(6144, 4096, 1, 0, "qkv")
run(4096, 4096, 1, 2, "o_proj")
