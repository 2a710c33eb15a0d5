#!/bin/bash
# L2 prefetch plan sweep (in-chain, whole model)
mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms')
"; }
ZL_L2_PREFETCH_MB=0 run "off"
ZL_L2_PREFETCH_MB=64 ZL_L2_PREFETCH_PLAN=12034 run "plan 12034 cap64"
ZL_L2_PREFETCH_MB=64 ZL_L2_PREFETCH_PLAN=02034 run "plan 02034 cap64"
ZL_L2_PREFETCH_MB=64 ZL_L2_PREFETCH_PLAN=02000 run "plan 02000 cap64 (attn->gu only)"
ZL_L2_PREFETCH_MB=32 ZL_L2_PREFETCH_PLAN=02000 run "plan 02000 cap32"
ZL_L2_PREFETCH_MB=64 ZL_L2_PREFETCH_PLAN=02030 run "plan 02030 cap64"
ZL_L2_PREFETCH_MB=64 ZL_L2_PREFETCH_PLAN=12334 run "plan 12334 cap64 (o->down too)"
ZL_L2_PREFETCH_MB=64 ZL_L2_PREFETCH_PLAN=20340 run "plan 20340 (qkv->gu, o->down, gu->nextqkv)"
