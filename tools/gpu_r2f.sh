#!/bin/bash
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
echo "== goldens from the reference library (for ref_kv8.npz)"; timeout 600 python -m oracle.gen_ref_golden $O/golden > $O/golden.log 2>&1; tail -2 $O/golden.log
echo "== dropin test"; timeout 900 python -m pytest tests/test_dropin_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
summ='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(d["value"],1), round(d["ms_per_step"],4), "step_roof", round(d["step_roofline"]["frac"],3), "gemm", round(d["roofline"]["us_per_launch"],2))
'
echo "== A/B: round-1 tree"; (cd _ab_r1 && timeout 600 python bench.py --steps 64 --warmup 4 2>/dev/null | python -c "$summ")
echo "== A/B: current tree"; timeout 600 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --no-extras --requests 0 2>/dev/null | python -c "$summ"
echo "== A/B: round-1 tree again"; (cd _ab_r1 && timeout 600 python bench.py --steps 64 --warmup 4 2>/dev/null | python -c "$summ")
echo "== A/B: current tree, prefill chunk 32"; timeout 600 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --no-extras --requests 0 --prefill-chunk 32 2>/dev/null | python -c "$summ"
echo "== tc bench (graph) after the polling fix"; timeout 600 python tools/tc_bench.py --graph > $O/tc_bench_graph.jsonl 2>$O/tc_bench_graph.err; python - <<'PY'
import json
for l in open('gpurun_out/r2f/tc_bench_graph.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
tail -3 $O/tc_bench_graph.err
for extra in "--batch 32"; do
echo "== bench $extra"; timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras --requests 0 $extra 2>/dev/null | python -c "$summ"
done
