#!/bin/bash
# Attribute the decode-step time by removing one kernel class at a time (ZL_DEBUG_SKIP), plus a test pass.
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest.log
run() { timeout 600 python bench.py --steps 64 --warmup 4 --no-cpu-baseline $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', {k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','kernels_per_step')}, 'gemm us/launch', round(d['roofline']['us_per_launch'],2), 'frac', round(d['roofline']['frac'],3))
"; }
run full ""
for m in 1 2 4 8 16 32 30 31 63; do ZL_DEBUG_SKIP=$m run "skip=$m" ""; done
run batch8 "--batch 8"
run batch32 "--batch 32"
