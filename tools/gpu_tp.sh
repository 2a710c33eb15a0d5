#!/bin/bash
# run with: gpurun --gpus N -- 'bash tools/gpu_tp.sh N'
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/tp_topo.txt 2>&1
timeout 600 python -m pytest tests/test_tp_gpu.py -m gpu -q -x --timeout 300 --timeout-method thread -p no:cacheprovider 2>&1 | tail -8
for n in 2 4 8; do
  if [ $n -le $N ]; then
    for extra in "" "--tp-int8"; do
      echo "== TP=$n $extra"
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 64 --warmup 4 --no-cpu-baseline $extra 2>&1 | grep '^{' | tee -a gpurun_out/tp_bench.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'],'gpus', round(d['value'],1),'tok/s', round(d['ms_per_step'],4),'ms e2e', round(d['e2e']['value'],1), d['config'].get('parallelism'))
"
    done
  fi
done
