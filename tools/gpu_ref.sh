#!/bin/bash
# Reference kernels on the box: goldens (pin the oracle), direct parity tests, reference-vs-ours timings.
mkdir -p gpurun_out/golden
timeout 600 python -m oracle.gen_ref_golden gpurun_out/golden > gpurun_out/gen_golden.log 2>&1; echo "golden rc=$?"; tail -3 gpurun_out/gen_golden.log
timeout 900 python -m pytest tests/test_vs_reference_gpu.py -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_ref.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_ref.log
timeout 600 python tools/ref_bench.py > gpurun_out/ref_bench.jsonl 2> gpurun_out/ref_bench.err; echo "ref_bench rc=$?"; cat gpurun_out/ref_bench.jsonl; tail -3 gpurun_out/ref_bench.err
