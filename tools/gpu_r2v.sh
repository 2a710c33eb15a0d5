#!/bin/bash
# composition of TTFT(128): ZL_DEBUG_SKIP drops kernels of the step (results are wrong, timing only)
mkdir -p gpurun_out/r2v
for skip in 0 1 30 31 32; do
echo "== ZL_DEBUG_SKIP=$skip (1 attention, 2 qkv, 4 o, 8 gate_up, 16 down, 32 lm_head)"
ZL_DEBUG_SKIP=$skip timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --requests 12 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d.get('latency'))
"
done
