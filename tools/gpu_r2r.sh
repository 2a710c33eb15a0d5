#!/bin/bash
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
echo "== tc + llama + layers tests"; timeout 1200 python -m pytest tests/test_w4a16_tc_gpu.py tests/test_llama_gpu.py tests/test_layers_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread -p no:cacheprovider 2>&1 | tail -4
echo "== tc bench (graph)"; timeout 600 python tools/tc_bench.py --graph > $O/tc_bench.jsonl 2>$O/tc_bench.err; python - <<PY
import json
for l in open('gpurun_out/r2r/tc_bench.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
tail -3 $O/tc_bench.err
echo "== tc bench (graph), ZL_TC_NO_PAIR=1, M=128/256 only"; ZL_TC_NO_PAIR=1 timeout 600 python tools/tc_bench.py --graph 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d['m']>=128: print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['tflops'],1),'TF')
"
summ='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(d["value"],1), round(d["ms_per_step"],4), "step_roof", round(d["step_roofline"]["frac"],3), d.get("logits_finite"), d.get("latency"))
'
echo "== bench latency (ttft), chunk 128"; timeout 900 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-extras 2>/dev/null | python -c "$summ"
echo "== bench latency (ttft), prompt 512 chunk 256"; timeout 900 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-extras --prompt 512 --prefill-chunk 256 2>/dev/null | python -c "$summ"
echo "== bench --batch 128"; timeout 900 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-extras --requests 0 --batch 128 2>&1 | python -c "$summ" | tail -3
