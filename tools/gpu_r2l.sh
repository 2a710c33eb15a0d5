#!/bin/bash
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
echo "== tc tests cfg 1 (4 groups x 4 warps)"; ZL_TC_CFG=1 timeout 900 python -m pytest tests/test_w4a16_tc_gpu.py -m gpu -x -q --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -4
for cfg in 1 0; do
echo "== tc bench (graph), ZL_TC_CFG=$cfg"; ZL_TC_CFG=$cfg timeout 600 python tools/tc_bench.py --graph > $O/tc_bench_cfg$cfg.jsonl 2>$O/tc_bench.err; python - <<PY
import json
for l in open('gpurun_out/r2l/tc_bench_cfg$cfg.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
tail -3 $O/tc_bench.err
done
echo "== trace 28672x4096 M=32 cfg 0"; timeout 200 python tools/tc_trace.py 28672 4096 32 > $O/trace_cfg0.txt 2>&1; sed -n 1,40p $O/trace_cfg0.txt
echo "== trace 28672x4096 M=32 cfg 1"; ZL_TC_CFG=1 timeout 200 python tools/tc_trace.py 28672 4096 32 > $O/trace_cfg1.txt 2>&1; sed -n 1,40p $O/trace_cfg1.txt
echo "== trace cfg 1, all ablated"; ZL_TC_CFG=1 ZL_TC_DBG=15 timeout 200 python tools/tc_trace.py 28672 4096 32 > $O/trace_cfg1_empty.txt 2>&1; sed -n 1,30p $O/trace_cfg1_empty.txt
echo "== trace cfg 1, M=128"; ZL_TC_CFG=1 timeout 200 python tools/tc_trace.py 28672 4096 128 > $O/trace_cfg1_m128.txt 2>&1; sed -n 1,24p $O/trace_cfg1_m128.txt
