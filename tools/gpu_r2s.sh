#!/bin/bash
# 2 GPUs: PDL policy under the fused tensor-parallel exchange
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
for mask in 70 64 66 68 6; do
  echo "== bench N=2 ZL_NO_PDL_MASK=$mask"
  ZL_NO_PDL_MASK=$mask timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 4 --no-cpu-baseline --requests 0 --no-tp-parity 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(round(d['value'],1), round(d['ms_per_step'],4), d.get('logits_finite'))
"
done
