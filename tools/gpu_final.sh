#!/bin/bash
# round-end record: smoke, all GPU tests, default bench, ncu launch list + full capture of the dominant kernel
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 64 --warmup 4 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 3 2>/dev/null | tail -1 | cut -c1-600
TAG=r01d
KREG='k_w4a16|k_add_rmsnorm|k_decode_attn|k_qkv_rope|k_dense|k_argmax|k_embedding|k_attn_combine|k_rope_cos|k_lens|k_advance|k_gate'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KREG" -s 1500 -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --prompt 2 --no-cpu-baseline --requests 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_w4a16 -s 300 -c 4 -f -o gpurun_out/${TAG}_prof_w4a16 python bench.py --steps 2 --warmup 3 --prompt 2 --no-cpu-baseline --requests 0 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
ZL_W4_DEBUG=10 LAYERS=8 timeout 200 python tools/trace_step.py > gpurun_out/r01_inchain_timeline.txt 2>&1; tail -6 gpurun_out/r01_inchain_timeline.txt
