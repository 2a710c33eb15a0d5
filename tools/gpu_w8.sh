#!/bin/bash
# W8 rows + layers mirror + reference goldens + in-chain timeline artefacts.
mkdir -p gpurun_out/golden
timeout 600 python -m oracle.gen_ref_golden gpurun_out/golden > gpurun_out/gen_golden.log 2>&1; echo "golden rc=$?"; tail -3 gpurun_out/gen_golden.log
timeout 300 python -m oracle.check_golden gpurun_out/golden 2>&1 | tail -15
timeout 900 python -m pytest tests/test_w8_gpu.py tests/test_layers_gpu.py tests/test_vs_reference_gpu.py -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_w8.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_w8.log
timeout 300 python tools/trace_step.py > gpurun_out/r01_inchain_timeline.txt 2>&1; echo "trace rc=$?"; tail -12 gpurun_out/r01_inchain_timeline.txt
timeout 120 tools/micro/mma_rate > gpurun_out/r01_mma_rate.txt 2>&1; echo "mma rc=$?"; cat gpurun_out/r01_mma_rate.txt | tail -8
timeout 300 python tools/w8_bench.py 2>&1 | tail -20
