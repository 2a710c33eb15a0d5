#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_llama_gpu.py -m gpu -q -x --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -12
timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
run() { timeout 300 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --requests 0 $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms', 'e2e', round(d['e2e']['value'],1), 'gemm us/launch', round(d['roofline']['us_per_launch'],2))
"; }
ZL_W4_TALL=1 run "tall-everywhere" ""
ZL_W4_TALL=2 run "wide-everywhere" ""
run "default" ""
timeout 300 python bench.py --model llama-3.2-1b --steps 64 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1500
