#!/bin/bash
# ncu evidence: per-launch durations of one decode step + a full capture of the dominant kernel.
mkdir -p gpurun_out
TAG=${1:-r01}
KREG='k_w4a16|k_add_rmsnorm|k_decode_attn|k_qkv_rope|k_dense|k_argmax|k_embedding|k_attn_combine|k_rope_cos|k_lens|k_advance|k_gate'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KREG" -s 1500 -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --prompt 2 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_w4a16 -s 300 -c 4 -f -o gpurun_out/${TAG}_prof_w4a16 python bench.py --steps 2 --warmup 3 --prompt 2 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | tail
