#!/bin/bash
# round 2, call A: tcgen05 bring-up probes + new parity tests + goldens + reference chain timing + regression run
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
echo "== probes"
(cd tools/micro && timeout 120 ./umma_i8_smoke > ../../$O/umma_i8.txt 2>&1; echo "i8 rc=$?"; timeout 120 ./umma_f16_probe > ../../$O/umma_f16.txt 2>&1; echo "f16 rc=$?")
cat $O/umma_i8.txt $O/umma_f16.txt
echo "== tc tests"
timeout 900 python -m pytest tests/test_w4a16_tc_gpu.py -q --timeout 240 --timeout-method thread -p no:cacheprovider > $O/pytest_tc.log 2>&1; echo "rc=$?"; tail -25 $O/pytest_tc.log
echo "== vs reference + int kernel tests"
timeout 900 python -m pytest tests/test_vs_reference_gpu.py tests/test_w4a16_int_gpu.py -q --timeout 240 --timeout-method thread -p no:cacheprovider > $O/pytest_ref.log 2>&1; echo "rc=$?"; tail -25 $O/pytest_ref.log
echo "== goldens"
timeout 600 python -m oracle.gen_ref_golden gpurun_out/golden > $O/gen_golden.log 2>&1; echo "rc=$?"; tail -3 $O/gen_golden.log
echo "== reference layer chain"
timeout 600 python tools/ref_layer_bench.py > $O/ref_layer.jsonl 2> $O/ref_layer.err; echo "rc=$?"; cat $O/ref_layer.jsonl; tail -3 $O/ref_layer.err
echo "== full gpu suite"
timeout 1800 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > $O/pytest_all.log 2>&1; echo "rc=$?"; tail -30 $O/pytest_all.log
for extra in "--batch 1" "--batch 16" "--batch 32"; do
echo "== bench $extra"; timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --requests 8 $extra 2>$O/bench_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k:d[k] for k in ('value','ms_per_step','kernels_per_step')}, 'e2e', round(d['e2e']['value'],1), 'gemm_roof', round(d['roofline']['frac'],3), 'step_roof', round(d['step_roofline']['frac'],3), d.get('latency'))
"; tail -3 $O/bench_err.txt
done
