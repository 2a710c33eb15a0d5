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
