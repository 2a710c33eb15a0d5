#!/bin/bash
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
echo "== tc tests"; timeout 900 python -m pytest tests/test_w4a16_tc_gpu.py -m gpu -x -q --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -4
echo "== tc bench (graph)"; timeout 600 python tools/tc_bench.py --graph > $O/tc_bench.jsonl 2>$O/tc_bench.err; python - <<PY
import json
for l in open('gpurun_out/r2m/tc_bench.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
tail -3 $O/tc_bench.err
echo "== trace 28672x4096 M=32"; timeout 200 python tools/tc_trace.py 28672 4096 32 > $O/trace_m32.txt 2>&1; sed -n 1,30p $O/trace_m32.txt
echo "== trace 4096x4096 M=32"; timeout 200 python tools/tc_trace.py 4096 4096 32 > $O/trace_small_m32.txt 2>&1; sed -n 1,10p $O/trace_small_m32.txt
echo "== ablations, graph mode"
for shape in "28672 4096 32" "4096 4096 32"; do
for dbg in 0 2 4 15; do
ZL_TC_DBG=$dbg timeout 120 python tools/tc_bench.py --one $shape --graph 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('shape $shape dbg $dbg', round(d['us'],1),'us', round(d['hbm_frac'],3))
"
done
done
echo "== grid sizes small GEMM"
for c in 32 64 96 128; do ZL_TC_CTAS=$c timeout 120 python tools/tc_bench.py --one 4096 4096 32 --graph 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ctas $c small', round(d['us'],1),'us', round(d['hbm_frac'],3))
"; done
summ='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(d["value"],1), round(d["ms_per_step"],4), "step_roof", round(d["step_roofline"]["frac"],3), "gemm", round(d["roofline"]["us_per_launch"],2), d.get("logits_finite"), d.get("latency"))
'
echo "== llama tests"; timeout 1200 python -m pytest tests/test_llama_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread -p no:cacheprovider 2>&1 | tail -3
for extra in "--batch 32"; do
echo "== bench $extra"; timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras --requests 0 $extra 2>/dev/null | python -c "$summ"
done
echo "== launch list of a batch-32 step (ncu, serialised)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 600 --csv --log-file $O/launches_b32.csv python bench.py --steps 2 --warmup 3 --batch 32 --no-cpu-baseline --no-extras --requests 0 --prompt 8 > $O/ncu_b32.log 2>&1; python tools/ncu_launch_summary.py $O/launches_b32.csv 2>&1 | head -20
