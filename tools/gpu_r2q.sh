#!/bin/bash
# 2 GPUs: tagged exchange with the cross-step parity rule
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
echo "== tp tests"
timeout 900 python -m pytest tests/test_tp_gpu.py -q -x --timeout 600 --timeout-method thread -p no:cacheprovider > $O/pytest_tp.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_tp.log
run_bench() {  # $1 = tag, rest = bench args
  tag=$1; shift
  echo "== bench N=2 $tag"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --requests 0 "$@" > $O/bench_n2_$tag.json 2> $O/bench_n2_$tag.err
  echo "rc=$?"; tail -1 $O/bench_n2_$tag.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k:d.get(k) for k in ('value','ms_per_step','kernels_per_step','logits_finite')}, 'parity', d.get('tp_parity',{}).get('ok'), d.get('tp_parity',{}).get('max_rel'), 'e2e', round(d['e2e']['value'],1), 'step_roof', round(d.get('step_roofline',{}).get('frac',0),3))
"; grep -v "OMP_NUM_THREADS\|^\*\*\*" $O/bench_n2_$tag.err | tail -3
}
run_bench 8b --steps 128 --warmup 4
run_bench 70b --model llama-3.1-70b --steps 32 --warmup 3 --no-tp-parity
run_bench 8b_b4 --steps 64 --warmup 4 --batch 4 --no-tp-parity
