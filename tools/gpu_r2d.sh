#!/bin/bash
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
echo "== nan check (32 layers, recentred synthetic weights)"; timeout 600 python tools/debug_nan.py 32 140 2 1 1 1 2>&1 | cut -c1-700
echo "== full gpu suite"
timeout 1800 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > $O/pytest_all.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_all.log
echo "== tc bench (graph)"; timeout 600 python tools/tc_bench.py --graph > $O/tc_bench_graph.jsonl 2>$O/tc_bench_graph.err; python - <<'PY'
import json
for l in open('gpurun_out/r2d/tc_bench_graph.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
tail -3 $O/tc_bench_graph.err
echo "== bench default (batch 1, extras)"; timeout 1500 python bench.py --steps 64 --warmup 4 > $O/bench_b1.json 2> $O/bench_b1.err; echo "rc=$?"; cat $O/bench_b1.json; tail -5 $O/bench_b1.err
for extra in "--batch 16" "--batch 32"; do
echo "== bench $extra"; timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras --requests 0 $extra 2>$O/bench_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k:d[k] for k in ('value','ms_per_step','kernels_per_step','logits_finite')}, 'e2e', round(d['e2e']['value'],1), 'gemm_roof', round(d['roofline']['frac'],3), d['roofline']['us_per_launch'], 'step_roof', round(d['step_roofline']['frac'],3))
"; tail -3 $O/bench_err.txt
done
echo "== attention / pdl variants at batch 1"
for v in "ZL_NO_PDL_MASK=70" "ZL_NO_PDL_MASK=64" "ZL_NO_PDL_MASK=68" "ZL_NO_PDL_MASK=66" "ZL_NO_PDL_MASK=70 ZL_ATTN_OLD_SHORT=1" "ZL_NO_PDL_MASK=64 ZL_ATTN_OLD_SHORT=1"; do
echo "-- $v"; env $v timeout 600 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --no-extras --requests 0 2>$O/var_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(d['value'],1), round(d['ms_per_step'],4), 'step_roof', round(d['step_roofline']['frac'],3), d['logits_finite'])
"; tail -2 $O/var_err.txt
done
echo "== in-chain timeline"; timeout 600 python tools/trace_step.py > $O/timeline.txt 2>&1; tail -6 $O/timeline.txt
echo "== attention bench"; timeout 600 python tools/attn_bench.py > $O/attn_bench.jsonl 2>$O/attn_bench.err; cat $O/attn_bench.jsonl; tail -3 $O/attn_bench.err
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 20 --warmup 3 --ref-budget 60 > $O/bench_ref.json 2>$O/bench_ref.err; cat $O/bench_ref.json; tail -2 $O/bench_ref.err
