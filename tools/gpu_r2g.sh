#!/bin/bash
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
echo "== tc tests (TS kernel)"; timeout 900 python -m pytest tests/test_w4a16_tc_gpu.py -m gpu -x -q --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -12
python - <<'PY'
import ctypes
from zhilight_b200 import _lib
print("watchdog code: %#x" % _lib.lib().zl_w4_tc_watchdog())
PY
summ='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(d["value"],1), round(d["ms_per_step"],4), "step_roof", round(d["step_roofline"]["frac"],3), "gemm", round(d["roofline"]["us_per_launch"],2), d.get("logits_finite"))
'
echo "== tc bench (graph), TS kernel"; timeout 600 python tools/tc_bench.py --graph > $O/tc_bench_ts.jsonl 2>$O/tc_bench_ts.err; python - <<'PY'
import json
for l in open('gpurun_out/r2g/tc_bench_ts.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
tail -3 $O/tc_bench_ts.err
echo "== A/B: round-1 tree"; (cd _ab_r1 && timeout 600 python bench.py --steps 64 --warmup 4 2>/dev/null | python -c "$summ")
echo "== A/B: current tree"; timeout 600 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --no-extras --requests 0 2>/dev/null | python -c "$summ"
echo "== llama tests"; timeout 1200 python -m pytest tests/test_llama_gpu.py tests/test_w4a16_int_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread -p no:cacheprovider 2>&1 | tail -5
for extra in "--batch 32" "--batch 16"; do
echo "== bench $extra"; timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras --requests 0 $extra 2>/dev/null | python -c "$summ"
done
echo "== bench latency (ttft)"; timeout 900 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d.get('latency'))
"
