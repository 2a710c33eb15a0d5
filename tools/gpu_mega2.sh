#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_llama_gpu.py -m gpu -q -x --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/trace_mega.py 2>&1 | tail -46
run() { timeout 300 python bench.py --steps 64 --warmup 4 --no-cpu-baseline $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms', 'e2e', round(d['e2e']['value'],1), 'kernels', d['kernels_per_step'])
"; }
run "fuse3" "--fuse 3"
run "fuse3 B=2" "--fuse 3 --batch 2"
run "fuse3 prompt=1000" "--fuse 3 --prompt 1000"
