#!/bin/bash
echo "== tc + llama tests"; timeout 1200 python -m pytest tests/test_w4a16_tc_gpu.py tests/test_llama_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread -p no:cacheprovider 2>&1 | tail -3
echo "== epilogue bench"; timeout 600 python tools/tc_epi_bench.py 2>&1 | grep -v '"none"' | tail -14
summ='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(d["value"],1), round(d["ms_per_step"],4), "step_roof", round(d["step_roofline"]["frac"],3), d.get("logits_finite"), d.get("latency"))
'
echo "== bench --batch 32"; timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras --requests 0 --batch 32 2>/dev/null | python -c "$summ"
echo "== bench latency"; timeout 900 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-extras --requests 16 2>/dev/null | python -c "$summ"
