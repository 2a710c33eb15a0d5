#!/usr/bin/env python
"""bench.py -- decode tokens/s of Llama-3.1-8B GPTQ-int4 (g128) on B200, the BASELINE.json metric.

    python bench.py --gpus N --steps K --warmup W [--batch B] [--prompt P] [--impl reference] [--model NAME]

A "step" is one decode step of the whole batch through the hot path (32 layers of W4A16 Linear +
decode attention + RMSNorm/RoPE fusions, lm_head, greedy pick).  Random-init weights of the named
architecture in HF-GPTQ layout, synthetic prompt token ids ("data": "synthetic").

  value   tokens/s with inputs resident in HBM: the step replays on the device, feeding the picked token
          back on the device (K steps between CUDA events on the driver's stream, max over ranks).
  e2e     the same step through the public host-buffer API (zl_llama_decode): pinned-host token ids and
          positions are copied H2D and the picked tokens are read back D2H inside every timed step.
  roofline  the dominant kernel (W4A16 GEMM): all 128 GEMM launches of a step replayed on their own real
          weights between CUDA events; achieved = algorithmic bytes / time against MEASURED_PEAKS.json.
  cpu_baseline  the oracle's dequant-to-bf16 torch CPU path on a bounded sample (1 of 32 layers + lm_head),
          a reported baseline, not a target.

--impl reference times that CPU path as the whole arm (the reference has no CPU implementation of its own;
SURVEY.md section 8d names this torch dequant path as the side-by-side baseline).

N > 1: tensor parallel over the N GPUs of one node (the reference's row/column split), one process per GPU; the
partial sums of the row-parallel Linears are exchanged with the one-shot NVLink peer-memory all-reduce
(zhilight_b200/csrc/comm.cu), NCCL only bootstraps.  value = tokens/s of the whole TP group ("scaling": "strong").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stamps = []          # host time of every sample
        self.t_region = None      # mark_region_start(): only samples taken after it describe the timed region
        self.stop_flag = False
        self.proc = None

    def mark_region_start(self):
        self.t_region = time.time()

    def _run_nvml(self):
        """5 ms NVML polling (nvidia_ml_py): enough samples inside a ~100 ms timed region.  Any failure falls back to
        the nvidia-smi loop below."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not self.stop_flag:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = int(get_reasons(h))
            flags = ["Active" if r & bits[k] else "Not Active"
                     for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]
            self.samples.append([str(self.gpu), str(sm), str(mx), "", ""] + flags)
            self.stamps.append(time.time())
            time.sleep(0.005)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
                self.stamps.append(time.time())
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        picked = [s for s, ts in zip(self.samples, self.stamps) if self.t_region is None or ts >= self.t_region]
        for s in (picked or self.samples):   # too short a region for the sampling period: fall back to all samples
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def model_cfg(name):
    from zhilight_b200.llama import MODEL_PRESETS
    return dict(MODEL_PRESETS[name])


def cpu_reference_path(cfg, quant, batch, budget_s=20.0, threads=None):
    """The CPU baseline: oracle dequant (W4 -> bf16) + matmul in torch on the host cores, on ONE layer's seven
    linears plus the lm_head, scaled to the layer count.  Returns (tokens/s, cores, sample description)."""
    import numpy as np
    import torch
    from oracle import gptq
    # torch's bf16 matmul stops scaling (and oversubscribes badly) beyond a few dozen threads
    threads = threads or min(os.cpu_count() or 1, 32)
    torch.set_num_threads(threads)
    d_model, d, ff, v = cfg["dim_model"], cfg["dim_head"], cfg["dim_ff"], cfg["vocab_size"]
    shapes = [(d_model, cfg["num_heads"] * d), (d_model, cfg["num_kv_heads"] * d), (d_model, cfg["num_kv_heads"] * d),
              (cfg["num_heads"] * d, d_model), (d_model, ff), (d_model, ff), (ff, d_model)]
    rng = np.random.default_rng(0)
    lin = []
    for k, n in shapes:
        if quant:
            qw = torch.from_numpy(rng.integers(0, 2 ** 31, size=(n, k // 8), dtype=np.int64).astype(np.int32))
            qz = torch.from_numpy(rng.integers(0, 16, size=(n, k // 128), dtype=np.int64).astype(np.uint8))
            sc = torch.from_numpy((0.002 + 0.004 * rng.random((n, k // 128))).astype(np.float32)).bfloat16()
            lin.append((qw, qz, sc, k, n))
        else:
            lin.append((torch.from_numpy(rng.standard_normal((n, k)).astype(np.float32)).bfloat16(), k, n))
    lm = torch.from_numpy(rng.standard_normal((v, d_model)).astype(np.float32) * 0.02).bfloat16()
    shifts = torch.tensor([gptq_shift(i) for i in range(8)], dtype=torch.int32)

    def one_pass():
        for item in lin:
            if quant:
                qw, qz, sc, k, n = item
                x = torch.randn(batch, k).bfloat16()
                # k-major unpack (q_gemm_k_major.cu:74-98 nibble order), (q - z) * s, to bf16 once per call
                q = ((qw.unsqueeze(-1) >> shifts) & 0xF).reshape(n, k).to(torch.bfloat16)
                w = (q - qz.to(torch.bfloat16).repeat_interleave(128, dim=1)) * sc.repeat_interleave(128, dim=1)
                _ = x @ w.T
            else:
                w, k, n = item
                x = torch.randn(batch, k).bfloat16()
                _ = x @ w.T
        x = torch.randn(batch, d_model).bfloat16()
        _ = x @ lm.T

    one_pass()
    t0 = time.perf_counter()
    n_pass = 0
    while True:
        one_pass()
        n_pass += 1
        if time.perf_counter() - t0 > budget_s or n_pass >= 8:
            break
    t_pass = (time.perf_counter() - t0) / n_pass
    # one pass = 1 layer + lm_head; a token needs num_layers layers + 1 lm_head.  Time the lm_head part separately.
    t1 = time.perf_counter()
    x = torch.randn(batch, d_model).bfloat16()
    _ = x @ lm.T
    t_lm = time.perf_counter() - t1
    t_layer = max(t_pass - t_lm, 1e-9)
    t_token_step = cfg["num_layers"] * t_layer + t_lm
    sample = ("1 of %d layers (7 linears, W4->bf16 dequant + matmul, M=%d) + lm_head, %d passes, scaled x%d"
              % (cfg["num_layers"], batch, n_pass, cfg["num_layers"]))
    return batch / t_token_step, threads, sample, t_token_step


def gptq_shift(kk):
    return (kk >> 1) * 4 + (kk & 1) * 16


def run_reference_arm(args, cfg, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    quant = args.model != "llama-3.2-1b"
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        v, cores, sample, t_step = cpu_reference_path(cfg, quant, args.batch, budget_s=15.0)
        vals.append((v, t_step))
    v = sum(x[0] for x in vals) / len(vals)
    t_step = sum(x[1] for x in vals) / len(vals)
    out = {
        "impl": "reference", "metric": "decode tokens/s", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload},
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--prompt", type=int, default=128)
    ap.add_argument("--model", default="llama-3.1-8b", choices=["llama-3.1-8b", "llama-3.2-1b", "tiny"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tp-int8", action="store_true", help="int8 group-32 payload for the TP all-reduce")
    ap.add_argument("--requests", type=int, default=32, help="requests for the TTFT/TPOT p50 (0 = skip)")
    ap.add_argument("--quant", choices=["default", "int8", "fp8"], default="default",
                    help="default: the model's BASELINE quantisation (GPTQ int4 for 8B, bf16 for 1B); int8 / fp8: W8A8 Linear rows")
    ap.add_argument("--prefill-chunk", type=int, default=128, help="tokens per chunked-prefill pass of the TTFT measurement")
    ap.add_argument("--fuse", type=int, default=2, help="0: one kernel per reference op; 1: +RMSNorm fused; 2: +qkv RoPE/KV epilogue")
    args = ap.parse_args()

    cfg = model_cfg(args.model)
    quant_name = "bf16" if args.model == "llama-3.2-1b" else "gptq-int4-g128-sym"
    if args.quant != "default":
        quant_name = {"int8": "w8a8-int8 (AutoInt8)", "fp8": "w8a8-fp8-e4m3"}[args.quant]
    workload = "%s %s greedy decode, batch=%d, prompt=%d, new=%d" % (args.model, quant_name, args.batch, args.prompt,
                                                                     args.steps)
    if args.impl == "reference":
        run_reference_arm(args, cfg, workload)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from zhilight_b200 import build
    from zhilight_b200.llama import LlamaDecoder
    build.build()

    from zhilight_b200 import dist as zdist
    dense = args.model == "llama-3.2-1b" or args.quant != "default"   # every non-W4 Linear path reports k_dense/k_w8a8
    qtype = {"default": 0 if args.model == "llama-3.2-1b" else 5, "int8": 2, "fp8": 7}[args.quant]
    W = max(args.warmup, 3)
    max_seq = args.prompt + 2 * W + 2 * args.steps + 16
    dec = LlamaDecoder(quant_type=qtype, group_size=128, sym=True, dtype="bf16" if args.model == "llama-3.2-1b" else "f16",
                       max_batch=args.batch, max_seq=max_seq, use_pdl=not args.no_pdl, use_graph=not args.no_graph, fuse=args.fuse,
                       tp_rank=rank, tp_size=world, tp_int8=args.tp_int8,
                       prefill_chunk=args.prefill_chunk if (world == 1 and args.requests > 0) else 0, **cfg)
    comm = None
    if world > 1:
        comm = zdist.TPComm(args.batch * cfg["dim_model"], rank, world)
        dec.set_comm(comm)
    dec.init_synthetic(seed=1)
    B = args.batch
    rng = np.random.default_rng(0)
    stream = torch.cuda.ExternalStream(dec.stream())

    # synthetic prompt ingested through the same decode path (fills the KV buffers to `prompt` tokens)
    prompt = rng.integers(0, cfg["vocab_size"], size=(args.prompt, B)).astype(np.int32)
    pos = np.zeros(B, dtype=np.int32)
    for t in range(args.prompt):
        dec.set_state(prompt[t], pos)
        dec.step_device(B)
        pos += 1
    dec.sync()
    tok, pos = dec.get_state(B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident loop ----
    for _ in range(W):
        dec.step_device(B)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)           # lets the nvidia-smi fallback start up; the NVML sampler is already polling
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = dec.lib.zl_launch_count(0)
    sampler.mark_region_start()
    e0.record(stream)
    for _ in range(args.steps):
        dec.step_device(B)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    clocks = sampler.finish()
    weight_bytes, kernels_per_step = dec.stats(B)

    # ---- e2e: host buffers in/out every step ----
    tok, pos = dec.get_state(B)
    for _ in range(W):
        tok = dec.decode(tok, pos)
        pos = pos + 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tok = dec.decode(tok, pos)
        pos = pos + 1
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3

    if world > 1:
        tt = torch.tensor([ms_dev, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = tt.tolist()

    # ---- roofline of the dominant kernel (W4A16 GEMM), live ----
    iters = 5
    g_ms, g_launches, g_bytes = dec.bench_gemms(B, iters)
    peak, peak_kind = read_peaks()
    achieved = g_bytes * iters / (g_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_w4a16_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch") if (not dense and B == 1 and world == 1) else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": ("k_w8a8_skinny (+ activation quant)" if args.quant != "default" else "k_dense_skinny") if dense else "k_w4a16_v3 (integer IMMA; k_w4a16_v2 where its staging does not fit)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind,
                "bytes_per_launch": g_bytes / g_launches, "us_per_launch": g_ms * 1e3 / iters / g_launches,
                "traffic": traffic}

    tokens = B * args.steps          # one TP group decodes B sequences, whatever its size
    value = tokens / (ms_dev * 1e-3)
    e2e_value = tokens / (ms_e2e * 1e-3)
    ctx_mid = args.prompt + W + args.steps // 2
    kv_bytes = B * ctx_mid * cfg["num_layers"] * 2 * cfg["num_kv_heads"] * cfg["dim_head"] * 2
    step_bytes = weight_bytes + kv_bytes
    step_roof = step_bytes / (ms_dev / args.steps * 1e-3) / 1e9

    out = {
        "metric": "decode tokens/s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": {"int8": "int8 x int8 -> int32 (W8A8), f16 activations", "fp8": "e4m3 x e4m3 -> f32 (W8A8), f16 activations"}.get(
            args.quant, "bf16" if dense else "f16 (W4A16: int4 weights, fp16 activations, fp32 accumulate)"),
        "data": "synthetic",
        "config": {"workload": workload, "parallelism": "single GPU" if world == 1 else "tp%d (one-shot NVLink all-reduce, %s payload)" % (world, "int8-g32" if args.tp_int8 else "fp16"),
                   "l2": "weights per step (%.2f GB) exceed L2 (126 MB); no explicit flush" % (weight_bytes / 1e9),
                   "pdl": not args.no_pdl, "cuda_graph": not args.no_graph, "fuse": args.fuse},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": 8 * B, "d2h_bytes_per_step": 4 * B,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(kernels_per_step) * args.steps if kernels_per_step else int(dec.lib.zl_launch_count(0) - launches0),
        "kernels_per_step": kernels_per_step,
        "roofline": roofline,
        "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "achieved": step_roof, "peak": peak, "unit": "GB/s",
                          "frac": step_roof / peak},
    }
    # ---- latency view of the same path: TTFT = chunked prefill of a fresh prompt + first token,
    # TPOT = (t_total - TTFT) / (n_out - 1), both wall clock through the host API, p50 over --requests requests ----
    if world == 1 and args.requests > 0:
        n_out = 16
        plen = min(args.prompt, max_seq - n_out - 1)
        ttft, tpot = [], []
        for r in range(args.requests):
            ptoks = rng.integers(0, cfg["vocab_size"], size=plen).astype(np.int32)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            t_next = dec.prefill(0, ptoks)
            t1 = time.perf_counter()
            tk = np.array([t_next], dtype=np.int32)
            ps = np.array([plen], dtype=np.int32)
            for _ in range(n_out - 1):
                tk = dec.decode(tk, ps)
                ps = ps + 1
            t2 = time.perf_counter()
            ttft.append((t1 - t0) * 1e3)
            tpot.append((t2 - t1) * 1e3 / (n_out - 1))
        out["latency"] = {"ttft_ms_p50": float(np.median(ttft)), "tpot_ms_p50": float(np.median(tpot)),
                          "requests": args.requests, "prompt_tokens": plen, "new_tokens": n_out, "batch": 1,
                          "prefill": "chunked, %d tokens per pass" % args.prefill_chunk}
    if rank == 0 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_path(cfg, not dense, B, budget_s=12.0)
        out["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(out))
    dec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
