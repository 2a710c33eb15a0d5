#!/bin/bash
# round 2, call B: fixes after call A, large-batch decode debugging, W8 tcgen05, int8-KV, bench
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
echo "== large batch decode (tcgen05 in the graph)"
timeout 900 python -m pytest tests/test_llama_gpu.py -q -k "large_batch or chunked_prefill or dual_stream or qkv_bias" --timeout 300 --timeout-method thread -p no:cacheprovider > $O/pytest_lb.log 2>&1; echo "rc=$?"; tail -30 $O/pytest_lb.log
echo "== tc + w8 + vs reference + attention"
timeout 1200 python -m pytest tests/test_w4a16_tc_gpu.py tests/test_w8_gpu.py tests/test_vs_reference_gpu.py tests/test_attention_gpu.py tests/test_w4a16_int_gpu.py tests/test_dropin_gpu.py -q --timeout 300 --timeout-method thread -p no:cacheprovider > $O/pytest_sel.log 2>&1; echo "rc=$?"; tail -30 $O/pytest_sel.log
echo "== goldens"
timeout 600 python -m oracle.gen_ref_golden gpurun_out/golden > $O/gen_golden.log 2>&1; echo "rc=$?"; tail -2 $O/gen_golden.log
echo "== full gpu suite"
timeout 1800 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > $O/pytest_all.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_all.log
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "rc=$?"; tail -3 $O/smoke.log
echo "== bench default (batch 1, extras)"; timeout 1200 python bench.py --steps 64 --warmup 4 > $O/bench_b1.json 2> $O/bench_b1.err; echo "rc=$?"; cat $O/bench_b1.json; tail -5 $O/bench_b1.err
for extra in "--batch 16" "--batch 32" "--batch 1 --no-extras --quant int8" "--batch 1 --no-extras --quant fp8"; do
echo "== bench $extra"; timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras --requests 0 $extra 2>$O/bench_err.txt | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k:d[k] for k in ('value','ms_per_step','kernels_per_step')}, 'e2e', round(d['e2e']['value'],1), 'gemm_roof', round(d['roofline']['frac'],3), d['roofline']['us_per_launch'], 'step_roof', round(d['step_roofline']['frac'],3))
"; tail -3 $O/bench_err.txt
done
echo "== attention / pdl variants at batch 1"
for v in "ZL_NO_PDL_MASK=70" "ZL_NO_PDL_MASK=64" "ZL_NO_PDL_MASK=68" "ZL_NO_PDL_MASK=66" "ZL_NO_PDL_MASK=70 ZL_ATTN_OLD_SHORT=1"; do
echo "-- $v"; env $v timeout 600 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --no-extras --requests 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(d['value'],1), round(d['ms_per_step'],4), 'step_roof', round(d['step_roofline']['frac'],3))
"
done
echo "== in-chain timeline"; timeout 600 python tools/trace_step.py > $O/timeline.txt 2>&1; tail -12 $O/timeline.txt
echo "== w8 microbench"; timeout 600 python tools/w8_bench.py > $O/w8_bench.txt 2>&1; tail -20 $O/w8_bench.txt
echo "== tc microbench"; timeout 600 python tools/tc_bench.py > $O/tc_bench.jsonl 2>$O/tc_bench.err; cat $O/tc_bench.jsonl; tail -3 $O/tc_bench.err
