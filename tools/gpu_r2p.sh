#!/bin/bash
# final single-GPU evidence run: full suite, smoke, bench lines, ncu captures
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== full gpu suite"
timeout 1800 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -p no:cacheprovider > $O/pytest_all.log 2>&1; echo "rc=$?"; tail -6 $O/pytest_all.log
echo "== bench default (batch 1, extras)"; timeout 1500 python bench.py --steps 64 --warmup 4 > $O/bench_b1.json 2> $O/bench_b1.err; echo "rc=$?"; cut -c1-1500 $O/bench_b1.json; tail -3 $O/bench_b1.err
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 20 --warmup 3 --ref-budget 60 > $O/bench_ref.json 2>$O/bench_ref.err; cut -c1-600 $O/bench_ref.json; tail -2 $O/bench_ref.err
summ='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(d["value"],1), round(d["ms_per_step"],4), "step_roof", round(d["step_roofline"]["frac"],3), "gemm", round(d["roofline"]["us_per_launch"],2), round(d["roofline"]["frac"],3), d.get("logits_finite"))
'
for extra in "--batch 32" "--batch 16" "--batch 4"; do
echo "== bench $extra"; timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras --requests 0 $extra 2>/dev/null | tee -a $O/bench_batches.jsonl | python -c "$summ"
done
echo "== bench llama-3.1-70b N=1"; timeout 1200 python bench.py --model llama-3.1-70b --steps 16 --warmup 3 --no-cpu-baseline --no-extras --requests 0 2>$O/bench_70b.err | tee $O/bench_70b.json | python -c "$summ"; tail -2 $O/bench_70b.err
echo "== ncu full: tcgen05 kernel 28672x4096 M=32"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_w4a16_ts -s 3 -c 1 -o $O/ts_m32 -f python tools/tc_bench.py --one 28672 4096 32 > $O/ncu_ts.log 2>&1; tail -1 $O/ncu_ts.log
echo "== ncu launch list, batch 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv --log-file $O/launches_b1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --requests 0 --prompt 2 > $O/ncu_b1.log 2>&1; python tools/ncu_launch_summary.py $O/launches_b1.csv 2>&1 | head -14
echo "== tc bench (graph)"; timeout 600 python tools/tc_bench.py --graph > $O/tc_bench.jsonl 2>/dev/null; python - <<PY
import json
for l in open('gpurun_out/r2p/tc_bench.jsonl'):
    d=json.loads(l); print(d['n'],d['k'],d['m'],round(d['us'],1),'us',round(d['hbm_frac'],3),'hbm',round(d['tflops'],1),'TF')
PY
echo "== attention bench (tensor-core P.V)"; timeout 600 python tools/attn_bench.py > $O/attn_bench.jsonl 2>$O/attn_bench.err; python - <<PY
import json
for l in open('gpurun_out/r2p/attn_bench.jsonl'):
    d=json.loads(l); print(d['kernel'], 'B', d['batch'], 'ctx', d['ctx'], round(d['us'],1), 'us', round(d['hbm_frac'],3))
PY
tail -2 $O/attn_bench.err
echo "== attention bench (ZL_ATTN_PV_FP32=1: round-1 CUDA-core P.V, fp16 KV only)"; ZL_ATTN_PV_FP32=1 timeout 600 python tools/attn_bench.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if 'fp16' in d['kernel']: print(d['kernel'], 'B', d['batch'], 'ctx', d['ctx'], round(d['us'],1), 'us', round(d['hbm_frac'],3))
"
