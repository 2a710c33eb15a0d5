#!/bin/bash
# 2 GPUs: TP parity tests, bench N=2 fused vs stand-alone exchange
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
nvidia-smi topo -m > $O/topo.txt 2>&1
echo "== tp tests"
timeout 1200 python -m pytest tests/test_tp_gpu.py -q --timeout 900 --timeout-method thread -p no:cacheprovider > $O/pytest_tp.log 2>&1; echo "rc=$?"; tail -25 $O/pytest_tp.log
run_bench() {  # $1 = tag, rest = env
  tag=$1; shift
  echo "== bench N=2 $tag"
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 4 --no-cpu-baseline > $O/bench_n2_$tag.json 2> $O/bench_n2_$tag.err
  echo "rc=$?"; tail -1 $O/bench_n2_$tag.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k:d.get(k) for k in ('value','ms_per_step','kernels_per_step','logits_finite','tp_parity')}, 'e2e', round(d['e2e']['value'],1))
"; tail -3 $O/bench_n2_$tag.err
}
run_bench fused ZL_DUMMY=1
run_bench unfused ZL_TP_UNFUSED=1
run_bench fused_pdl64 ZL_NO_PDL_MASK=64
echo "== bench N=2 llama-3.1-70b"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --model llama-3.1-70b --steps 16 --warmup 3 --no-cpu-baseline --no-tp-parity > $O/bench_n2_70b.json 2> $O/bench_n2_70b.err; echo "rc=$?"; tail -1 $O/bench_n2_70b.json | cut -c1-900; tail -3 $O/bench_n2_70b.err
