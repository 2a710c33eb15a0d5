#!/bin/bash
mkdir -p gpurun_out/golden
timeout 600 python -m oracle.gen_ref_golden gpurun_out/golden > gpurun_out/gen_golden.log 2>&1; echo "golden rc=$?"; tail -3 gpurun_out/gen_golden.log
timeout 300 python -m oracle.check_golden gpurun_out/golden 2>&1 | tail -15
timeout 900 python -m pytest tests/test_vs_reference_gpu.py tests/test_attention_gpu.py tests/test_llama_gpu.py -m gpu -q -x --timeout 300 --timeout-method thread -p no:cacheprovider 2>&1 | tail -5
timeout 300 python tools/w8_bench.py 2>&1 | tail -26
run() { timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms')
"; }
ZL_L2_PREFETCH_MB=0 run "off"
export ZL_W4_DEBUG=4
ZL_L2_PREFETCH_MB=64 ZL_L2_PREFETCH_PLAN=12034 run "lines plan 12034 cap64"
ZL_L2_PREFETCH_MB=64 ZL_L2_PREFETCH_PLAN=02000 run "lines plan 02000 cap64 (attn->gu only)"
ZL_L2_PREFETCH_MB=32 ZL_L2_PREFETCH_PLAN=02000 run "lines plan 02000 cap32"
ZL_L2_PREFETCH_MB=64 ZL_L2_PREFETCH_PLAN=10234 run "lines plan 10234 cap64 (each GEMM -> next GEMM)"
ZL_L2_PREFETCH_MB=16 ZL_L2_PREFETCH_PLAN=10234 run "lines plan 10234 cap16"
ZL_L2_PREFETCH_MB=64 ZL_L2_PREFETCH_PLAN=12004 run "lines plan 12004 cap64"
