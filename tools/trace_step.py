#!/usr/bin/env python
"""In-chain timeline of the W4A16 launches of one decode step (ZL_W4_DEBUG=2 + zl_w4_set_trace)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ZL_W4_DEBUG"] = os.environ.get("ZL_W4_DEBUG", "2")   # 10 = 2 | 8: extra stamps inside staging
from zhilight_b200 import _lib  # noqa: E402
from zhilight_b200.llama import LlamaDecoder, MODEL_PRESETS  # noqa: E402

torch.cuda.set_device(0)
lib = _lib.load()
cfg = dict(MODEL_PRESETS["llama-3.1-8b"])
layers = int(os.environ.get("LAYERS", "8"))
cfg["num_layers"] = layers
dec = LlamaDecoder(quant_type=5, sym=True, max_batch=1, max_seq=512, **cfg)
dec.init_synthetic(1)
dec.set_state(np.array([1], np.int32), np.array([0], np.int32))
for _ in range(130):
    dec.step_device(1)
dec.sync()
trace = torch.zeros(8 + 16 * 4096, dtype=torch.int64, device="cuda")
lib.zl_w4_set_trace(ctypes.c_void_p(trace.data_ptr()))
# graphs captured earlier hold trace == NULL in their kernel params: force a re-capture by a new batch bucket
dec2 = LlamaDecoder(quant_type=5, sym=True, max_batch=1, max_seq=512, **cfg)
dec2.init_synthetic(1)
dec2.set_state(np.array([1], np.int32), np.array([140], np.int32))
for _ in range(3):
    dec2.step_device(1)
dec2.sync()
trace.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.ExternalStream(dec2.stream())
e0.record(st)
dec2.step_device(1)
e1.record(st)
dec2.step_device(1)     # second step: shows the step boundary (head of step 2 after the tail of step 1)
dec2.sync()
t = trace.cpu().numpy()
n = int(t[0])
rec = t[8:8 + n * 16].reshape(n, 16)
rec = rec[np.argsort(rec[:, 0])]
t0 = rec[0, 0]
dense = rec[rec[:, 15] == 0xD]
gemm_all = rec[rec[:, 15] != 0xD]
if len(dense):
    d = dense[0]
    before = gemm_all[gemm_all[:, 0] < d[0]]
    after = gemm_all[gemm_all[:, 2] > d[3]]
    last_end = max(x for x in before[-1] if x > 1000) if len(before) else d[0]
    print("TAIL of step 1: last layer GEMM end -> lm_head entry %+.2f us, wait_done %+.2f, lm_head CTA0 end %+.2f ; "
          "next step first qkv wait_done %+.2f us after lm_head CTA0 end"
          % ((d[0] - last_end) / 1e3, (d[2] - last_end) / 1e3, (d[3] - last_end) / 1e3,
             ((after[0][2] - d[3]) / 1e3) if len(after) else float("nan")))
if len(dense):
    hd = t[:8]
    print("      lm_head last CTA done %+.2f us after last GEMM end; k_advance (after the graph) %+.2f ; next step k_lens_from_pos %+.2f"
          % ((hd[5] - last_end) / 1e3, (hd[2] - last_end) / 1e3, (hd[1] - last_end) / 1e3))
rec = gemm_all[: 4 * layers]
n = len(rec)
print("step %.1f us, %d GEMM launches, %d layers" % (e0.elapsed_time(e1) * 1e3, n, layers))
names = ["qkv", "o", "gate_up", "down"]
prev_end = 0.0
for i, r in enumerate(rec[: 4 * min(layers, 4)]):
    v = [(x - t0) / 1e3 for x in r if x > 0]
    end = v[-1]
    print("%-8s entry %7.2f  ring %6.2f  wait_done %7.2f  staged %7.2f  end %7.2f | span %5.2f  since_prev_end %6.2f  staging %5.2f" % (
        names[i % 4], v[0], v[1] - v[0], v[2], v[3], end, end - v[2], v[2] - prev_end, v[3] - v[2]))
    if int(os.environ["ZL_W4_DEBUG"]) & 8:
        # stamps: entry, ring, wait_done, warp0 staged, all staged, rstd ready (= "staged"), tiles..., end
        print("         raw deltas since wait_done: " + " ".join("%.2f" % (x - v[2]) for x in v[3:]))
    prev_end = end
per = {k: [] for k in names}
gaps = {k: [] for k in names}
prev_end = None
for i, r in enumerate(rec):
    v = [x for x in r if x > 0]
    per[names[i % 4]].append((v[-1] - v[2]) / 1e3)
    if prev_end is not None:
        gaps[names[i % 4]].append((v[2] - prev_end) / 1e3)
    prev_end = v[-1]
for k in names:
    print("%-8s CTA0 active span (wait_done->end) median %.2f us ; gap since previous GEMM's end median %.2f us" % (
        k, np.median(per[k]), np.median(gaps[k]) if gaps[k] else float("nan")))
