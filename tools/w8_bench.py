#!/usr/bin/env python
"""Isolated timing of the W8A8 Linear kernels (quant + fused GEMM) on Llama-3.1-8B layer shapes, weights rotated
through > L2 worth of copies; prints JSON lines with GB/s against the W8 algorithmic bytes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zhilight_b200 import ops  # noqa: E402

PEAK = 6567.7
dev = torch.device("cuda:0")
for kind in ("int8", "fp8"):
    for n, k in ((6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336)):
        copies = max(2, int(300e6 // (n * k)) + 1)
        ws = [torch.randint(-127, 128, (n, k), dtype=torch.int8, device=dev) for _ in range(copies)]
        if kind == "fp8":
            ws = [(w.view(torch.uint8) & 0x77) for w in ws]
        for m in (1, 8, 32):
            x = torch.randn(m, k, device=dev).half()
            sw = torch.full((n,), 0.01, dtype=torch.float16, device=dev) if kind == "int8" else torch.tensor([0.01], device=dev)
            f = ops.int8_linear if kind == "int8" else ops.fp8_linear
            for i in range(copies):
                f(x, ws[i], sw)
            torch.cuda.synchronize()
            # one CUDA graph over `iters` back-to-back Linears: no host launch overhead in the number
            iters = 4 * copies
            st = torch.cuda.Stream()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(st):
                with torch.cuda.graph(g, stream=st):
                    for i in range(iters):
                        f(x, ws[i % copies], sw, pdl=True)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            nbytes = n * k + n * 2 + m * k + m * 4 + m * n * 2
            print(json.dumps({"kernel": "w8a8_" + kind, "n": n, "k": k, "m": m, "us_quant_plus_gemm": round(us, 2),
                              "gbs": round(nbytes / us / 1e3, 1), "frac": round(nbytes / us / 1e3 / PEAK, 3)}))
        del ws
