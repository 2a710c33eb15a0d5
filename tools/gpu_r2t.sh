#!/bin/bash
# final evidence run on the committed tree
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full gpu suite"
timeout 1800 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > $O/pytest_all.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_all.log
echo "== bench default (batch 1, extras)"; timeout 1500 python bench.py --steps 64 --warmup 4 > $O/bench_b1.json 2> $O/bench_b1.err; echo "rc=$?"; cut -c1-400 $O/bench_b1.json; tail -3 $O/bench_b1.err
echo "== ncu launch list of decode steps, batch 1 (step kernels only)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_w4a16_v3|k_decode_attn|k_dense_skinny|k_add_rmsnorm|k_rmsnorm|k_argmax|k_embedding|k_rope_cos_sin|k_lens_from_pos|k_advance|k_attn_combine" -s 700 -c 600 --csv --log-file $O/launches_b1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --requests 0 --prompt 2 > $O/ncu_b1.log 2>&1; python tools/ncu_launch_summary.py $O/launches_b1.csv 2>&1 | head -14
echo "== tc bench (graph)"; timeout 600 python tools/tc_bench.py --graph > $O/tc_bench.jsonl 2>/dev/null; python - <<PY
import json
for l in open('gpurun_out/r2t/tc_bench.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
echo "== epilogue bench"; timeout 600 python tools/tc_epi_bench.py > $O/tc_epi_bench.jsonl 2>/dev/null; cat $O/tc_epi_bench.jsonl | tail -24
for extra in "--batch 32" "--batch 16"; do
echo "== bench $extra"; timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras --requests 0 $extra 2>/dev/null | tee -a $O/bench_batches.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(round(d['value'],1), round(d['ms_per_step'],4), 'step_roof', round(d['step_roofline']['frac'],3), d.get('logits_finite'))
"
done
