#!/usr/bin/env python
"""Pipeline hand-over trace of the tcgen05 W4A16 kernel (CTA 0): ZL_TC_DBG=16 makes every role stamp clock64 at its barrier
waits; this prints, per stage, when each role got what it waited for (cycles since the first stamp).  Engineering probe."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ZL_TC_DBG"] = str(int(os.environ.get("ZL_TC_DBG", "0")) | 16)
from zhilight_b200 import _lib, build, ops  # noqa: E402

build.build()
lib = _lib.load()
n, k, m = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (28672, 4096, 32)
dev = torch.device("cuda:0")
nbytes = (n // 32) * (k // 128) * 2128
pack = torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=dev)
v = pack.view(-1, 2128)
v[:, 2048:2112] = 0
v[:, 2049:2112:2] = 0x1c
x = torch.randn(m, k, device=dev).half()
out = torch.empty(m, n, device=dev).half()
for _ in range(3):
    ops.w4a16_gemm_fused(x, pack, n, k, out=out, variant=1)
torch.cuda.synchronize()
ROLES = 11
buf = np.zeros(ROLES * 64 * 8, np.int64)
lib.zl_w4_tc_read_trace.restype = ctypes.c_int
rc = lib.zl_w4_tc_read_trace(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), buf.size)
assert rc == 0, rc
t = buf.reshape(ROLES, 64, 8)
t0 = t[t > 0].min()
r = lambda a: int(a - t0) if a > 0 else -1
print("cycles since first stamp (a stage = 2 quantisation groups for <= 128 tokens).  raw: got raw_empty | x: got ax_empty | dq (warp 4): "
      "start, got raw_full, got ax_empty, arrived a_full | slowest dq arrive | mma: start, got a_full, issued + committed")
for s in range(32):
    dq_arr = max(t[3 + w, s, 3] for w in range(8))
    d = t[3, s]
    mm = t[2, s]
    print("st %2d | raw %6d | x %6d | dq %6d %6d %6d %6d | dqmax %6d | mma %6d %6d %6d" % (
        s, r(t[0, s, 1]), r(t[1, s, 1]), r(d[0]), r(d[1]), r(d[2]), r(d[3]), r(dq_arr), r(mm[0]), r(mm[1]), r(mm[3])))
