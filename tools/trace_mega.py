#!/usr/bin/env python
"""Phase timeline of the persistent decode kernel (CTA 0, %globaltimer): ZL_MEGA_TRACE=1 + zl_llama_mega_trace."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ZL_MEGA_TRACE"] = "1"
from zhilight_b200 import _lib  # noqa: E402
from zhilight_b200.llama import LlamaDecoder, MODEL_PRESETS  # noqa: E402

torch.cuda.set_device(0)
lib = _lib.load()
cfg = dict(MODEL_PRESETS["llama-3.1-8b"])
layers = int(os.environ.get("LAYERS", "16"))
cfg["num_layers"] = layers
dec = LlamaDecoder(quant_type=5, sym=True, max_batch=1, max_seq=512, fuse=3, **cfg)
dec.init_synthetic(1)
dec.set_state(np.array([1], np.int32), np.array([0], np.int32))
for _ in range(int(os.environ.get("CTX", "130"))):
    dec.step_device(1)
dec.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.ExternalStream(dec.stream())
e0.record(st)
dec.step_device(1)
e1.record(st)
dec.sync()
buf = (ctypes.c_ulonglong * 512)()
_lib.check(lib.zl_llama_mega_trace(dec.h, buf, 512))
t = np.frombuffer(buf, dtype=np.uint64).astype(np.int64)
n = int(t[0])
rec = t[1:1 + 2 * n].reshape(n, 2)
print("step %.1f us, %d records, %d layers" % (e0.elapsed_time(e1) * 1e3, n, layers))
NAMES = {0: "entry", 1: "wait_done", 10: "qkv computed", 20: " barrier", 30: "attn computed", 31: " barrier", 32: "combine+barrier",
         11: "o computed", 21: " barrier", 12: "gate_up computed", 22: " barrier", 13: "down computed", 23: " barrier"}
t0 = rec[0, 1]
prev = t0
per = {}
for i, (idn, ts) in enumerate(rec):
    dt = (ts - prev) / 1e3
    key = int(idn) % 1000
    if int(idn) >= 10000:
        ph, sub = int(idn) // 10000 - 1, int(idn) % 1000
        print("L%-2d   phase %d %-10s t=%9.2f us  (+%6.2f)" % ((int(idn) % 10000) // 1000, ph, "staged" if sub == 100 else "tile %d k-loop done" % (sub - 200), (ts - t0) / 1e3, dt))
        prev = ts
        continue
    if i < 60:
        print("L%-2d %-16s t=%9.2f us  (+%6.2f)" % (int(idn) // 1000, NAMES.get(key, str(key)), (ts - t0) / 1e3, dt))
    per.setdefault(key, []).append(dt)
    prev = ts
print("medians over layers (us since previous stamp):")
order = [10, 20, 30, 31, 32, 11, 21, 12, 22, 13, 23]
tot = 0.0
for k in order:
    if k in per:
        m = float(np.median(per[k][1:])) if len(per[k]) > 1 else per[k][0]
        tot += m
        print("  %-16s %7.2f" % (NAMES[k], m))
print("  per layer        %7.2f" % tot)
