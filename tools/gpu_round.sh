#!/bin/bash
# One gpurun call: smoke, GPU tests, bench variants.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest.log
echo "== bench default"; timeout 900 python bench.py --steps 64 --warmup 4 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for extra in "--fuse 1" "--fuse 0" "--fuse 2 --no-pdl" "--batch 8" "--batch 32"; do
echo "== bench $extra"; timeout 900 python bench.py --steps 64 --warmup 4 --no-cpu-baseline $extra 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k:d[k] for k in ('value','ms_per_step','kernels_per_step')}, 'e2e', round(d['e2e']['value'],1), 'gemm_roof', round(d['roofline']['frac'],3), 'us/launch', round(d['roofline']['us_per_launch'],2), 'step_roof', round(d['step_roofline']['frac'],3))
"
done
echo "== v1 kernel A/B"; ZL_W4_KERNEL=1 timeout 900 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --fuse 0 2>&1 | cut -c1-400
