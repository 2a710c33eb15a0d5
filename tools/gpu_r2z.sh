#!/bin/bash
echo "== tp tests"; timeout 600 python -m pytest tests/test_tp_gpu.py -q -x --timeout 500 --timeout-method thread -p no:cacheprovider 2>&1 | tail -3
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 4 --no-cpu-baseline --requests 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(round(d['value'],1), round(d['ms_per_step'],4), d.get('logits_finite'), d.get('tp_parity'))
"
