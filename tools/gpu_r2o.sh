#!/bin/bash
# 8 GPUs: scaling of the headline config and the 70B preset
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
nvidia-smi topo -m > $O/topo.txt 2>&1
run_bench() {  # $1 = tag, $2 = nproc, rest = bench args
  tag=$1; n=$2; shift; shift
  echo "== bench N=$n $tag"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --no-cpu-baseline --requests 0 "$@" > $O/bench_${tag}.json 2> $O/bench_${tag}.err
  echo "rc=$?"; tail -1 $O/bench_${tag}.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k:d.get(k) for k in ('value','ms_per_step','kernels_per_step','logits_finite')}, 'parity', d.get('tp_parity',{}).get('ok'), d.get('tp_parity',{}).get('max_rel'), 'e2e', round(d['e2e']['value'],1), 'step_roof', round(d.get('step_roofline',{}).get('frac',0),3))
"; grep -v "OMP_NUM_THREADS\|^\*\*\*" $O/bench_${tag}.err | tail -3
}
run_bench n8_8b 8 --steps 64 --warmup 4
run_bench n4_8b 4 --steps 64 --warmup 4
run_bench n8_70b 8 --model llama-3.1-70b --steps 32 --warmup 3 --no-tp-parity
run_bench n4_70b 4 --model llama-3.1-70b --steps 32 --warmup 3 --no-tp-parity
ZL_TP_UNFUSED=1 run_bench n8_8b_unfused 8 --steps 64 --warmup 4 --no-tp-parity
