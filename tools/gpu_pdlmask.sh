#!/bin/bash
run() { timeout 300 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --requests 0 $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms')
"; }
for mk in 64 0 65 66 68 72 80 96 192 194 224; do ZL_NO_PDL_MASK=$mk run "no-pdl mask $mk" ""; done
