#!/bin/bash
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
echo "== nan debug"; timeout 1500 python tools/debug_nan.py > $O/debug_nan.txt 2>&1; cat $O/debug_nan.txt | cut -c1-600
echo "== tests (previous failures)"
timeout 900 python -m pytest tests/test_w4a16_tc_gpu.py tests/test_dropin_gpu.py tests/test_vs_reference_gpu.py "tests/test_llama_gpu.py::test_dual_stream_prefill_matches_single_stream" -q --timeout 300 --timeout-method thread -p no:cacheprovider > $O/pytest_sel.log 2>&1; echo "rc=$?"; tail -12 $O/pytest_sel.log
echo "== tc bench (stream)"; timeout 600 python tools/tc_bench.py > $O/tc_bench.jsonl 2>$O/tc_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/r2c/tc_bench.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
echo "== tc bench (graph)"; timeout 600 python tools/tc_bench.py --graph > $O/tc_bench_graph.jsonl 2>$O/tc_bench_graph.err; python - <<'PY'
import json
for l in open('gpurun_out/r2c/tc_bench_graph.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
tail -3 $O/tc_bench_graph.err
echo "== ncu tc kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_w4a16_tc -s 6 -c 2 -o $O/tc_prof python tools/tc_bench.py --one 28672 4096 32 > $O/ncu_tc.log 2>&1; tail -3 $O/ncu_tc.log
echo "== attention bench"; timeout 600 python tools/attn_bench.py > $O/attn_bench.jsonl 2>$O/attn_bench.err; cat $O/attn_bench.jsonl; tail -3 $O/attn_bench.err
echo "== w8 bench"; timeout 600 python tools/w8_bench.py > $O/w8_bench.txt 2>&1; tail -24 $O/w8_bench.txt
