#!/bin/bash
# 2 GPUs: tagged (LL) exchange inside the GEMMs
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
echo "== tp tests"
timeout 900 python -m pytest tests/test_tp_gpu.py -q -x --timeout 600 --timeout-method thread -p no:cacheprovider > $O/pytest_tp.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_tp.log
run_bench() {  # $1 = tag, rest = env
  tag=$1; shift
  echo "== bench N=2 $tag"
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 4 --no-cpu-baseline --requests 0 > $O/bench_n2_$tag.json 2> $O/bench_n2_$tag.err
  echo "rc=$?"; tail -1 $O/bench_n2_$tag.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k:d.get(k) for k in ('value','ms_per_step','kernels_per_step','logits_finite')}, 'parity', d.get('tp_parity',{}).get('ok'), d.get('tp_parity',{}).get('max_rel'), 'e2e', round(d['e2e']['value'],1))
"; grep -v "OMP_NUM_THREADS\|^\*\*\*" $O/bench_n2_$tag.err | tail -3
}
run_bench fused ZL_DUMMY=1
run_bench unfused ZL_TP_UNFUSED=1
run_bench fused_b4 ZL_DUMMY=1 ZL_BENCH_BATCH=4
