#!/bin/bash
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
echo "== tc tests"; timeout 900 python -m pytest tests/test_w4a16_tc_gpu.py -m gpu -x -q --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -6
echo "== tc tests, one dequant group"; ZL_TC_GROUPS=1 timeout 900 python -m pytest tests/test_w4a16_tc_gpu.py -m gpu -x -q --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -3
for ng in 2 1; do
echo "== tc bench (graph), ZL_TC_GROUPS=$ng"; ZL_TC_GROUPS=$ng timeout 600 python tools/tc_bench.py --graph > $O/tc_bench_ng$ng.jsonl 2>$O/tc_bench.err; python - <<PY
import json
for l in open('gpurun_out/r2k/tc_bench_ng$ng.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
tail -3 $O/tc_bench.err
done
echo "== trace 28672x4096 M=32 (2 groups)"; timeout 200 python tools/tc_trace.py 28672 4096 32 > $O/trace_big32.txt 2>&1; sed -n 1,34p $O/trace_big32.txt
echo "== trace 28672x4096 M=32 (1 group)"; ZL_TC_GROUPS=1 timeout 200 python tools/tc_trace.py 28672 4096 32 > $O/trace_big32_ng1.txt 2>&1; sed -n 1,26p $O/trace_big32_ng1.txt
echo "== ablations, graph mode, 2 groups"
for shape in "28672 4096 32" "28672 4096 128"; do
for dbg in 0 1 2 4 8 15; do
ZL_TC_DBG=$dbg timeout 120 python tools/tc_bench.py --one $shape --graph 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('shape $shape dbg $dbg', round(d['us'],1),'us', round(d['hbm_frac'],3))
"
done
done
summ='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(d["value"],1), round(d["ms_per_step"],4), "step_roof", round(d["step_roofline"]["frac"],3), "gemm", round(d["roofline"]["us_per_launch"],2), d.get("logits_finite"), d.get("latency"))
'
echo "== llama tests"; timeout 1200 python -m pytest tests/test_llama_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread -p no:cacheprovider 2>&1 | tail -3
for extra in "--batch 32" "--batch 16"; do
echo "== bench $extra"; timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras --requests 0 $extra 2>/dev/null | python -c "$summ"
done
echo "== bench latency (ttft)"; timeout 900 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | python -c "$summ"
