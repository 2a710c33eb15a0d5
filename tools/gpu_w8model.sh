#!/bin/bash
timeout 600 python -m pytest tests/test_llama_gpu.py tests/test_w8_gpu.py -m gpu -q -x --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -4
run() { timeout 300 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --requests 8 $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms', 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'step_roof', round(d['step_roofline']['frac'],3), d.get('latency',{}).get('ttft_ms_p50'))
"; }
run "8B int8 B=1" "--quant int8"
run "8B fp8 B=1" "--quant fp8"
run "8B int8 B=8" "--quant int8 --batch 8"
