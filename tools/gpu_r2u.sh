#!/bin/bash
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
echo "== launch list of prefill passes (128-token prompt, one 128-token chunk): kernels that only prefill launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_w4a16_ts|k_decode_attn<|k_prefill_setup|k_attn_combine|k_qkv_rope" -c 1200 --csv --log-file $O/launches_prefill.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --requests 3 > $O/ncu_prefill.log 2>&1
python tools/ncu_launch_summary.py $O/launches_prefill.csv 2>&1 | head -24
tail -3 $O/ncu_prefill.log
