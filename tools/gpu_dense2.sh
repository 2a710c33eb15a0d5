#!/bin/bash
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_llama_gpu.py tests/test_layers_gpu.py -m gpu -q -x --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -3
ZL_W4_DEBUG=2 LAYERS=4 timeout 200 python tools/trace_step.py 2>&1 | grep -E "TAIL|lm_head last|^step"
run() { timeout 300 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --requests 0 $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms', 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'step_roof', round(d['step_roofline']['frac'],3))
"; }
run "8B B=1" ""
run "1B bf16" "--model llama-3.2-1b"
