#!/bin/bash
# One gpurun call: smoke, GPU tests, micro-benchmarks, bench, ncu launch list + one full capture.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest.log
echo "== micro"; timeout 600 python tools/microbench.py > gpurun_out/micro.jsonl 2> gpurun_out/micro.err; echo "rc=$?"; cat gpurun_out/micro.jsonl | cut -c1-200; tail -5 gpurun_out/micro.err
echo "== bench"; timeout 900 python bench.py --steps 64 --warmup 4 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_w4a16|k_add_rmsnorm|k_decode_attn|k_qkv_rope|k_dense|k_argmax|k_embedding|k_attn_combine|k_rope_cos|k_lens|k_advance' -s 1320 -c 560 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --prompt 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_w4a16 -s 300 -c 4 -o gpurun_out/prof_w4a16 -f python bench.py --steps 2 --warmup 3 --prompt 2 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
fi
