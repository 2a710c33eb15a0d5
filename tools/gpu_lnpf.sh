#!/bin/bash
timeout 600 python -m pytest tests/test_llama_gpu.py tests/test_w4a16_int_gpu.py -m gpu -q -x --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -2
run() { timeout 300 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --requests 0 $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms', 'e2e', round(d['e2e']['value'],1))
"; }
run "default (mask 70 + ln prefetch)" ""
ZL_NO_LN_PREFETCH=1 run "no ln prefetch" ""
ZL_NO_PDL_MASK=64 run "mask 64" ""
run "B=2" "--batch 2"
run "B=4" "--batch 4"
run "B=8" "--batch 8"
run "B=16" "--batch 16"
run "B=32" "--batch 32"
