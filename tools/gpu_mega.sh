#!/bin/bash
# persistent-kernel bring-up: every run wrapped in its own timeout (the kernel aborts itself after 2 s at a barrier)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_llama_gpu.py -m gpu -q -x --timeout 120 --timeout-method thread -p no:cacheprovider -k "persistent or fuse3 or 3-" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_llama_gpu.py tests/test_attention_gpu.py -m gpu -q -x --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -5
run() { timeout 300 python bench.py --steps 64 --warmup 4 --no-cpu-baseline $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', round(d['value'],1), 'tok/s', round(d['ms_per_step'],4), 'ms', 'e2e', round(d['e2e']['value'],1), 'kernels', d['kernels_per_step'])
"; }
run "fuse2" "--fuse 2"
run "fuse3" "--fuse 3"
run "fuse3 B=2" "--fuse 3 --batch 2"
run "fuse3 B=4" "--fuse 3 --batch 4"
run "fuse3 B=8" "--fuse 3 --batch 8"
run "fuse2 B=8" "--fuse 2 --batch 8"
run "fuse3 prompt=1000" "--fuse 3 --prompt 1000"
run "fuse2 prompt=1000" "--fuse 2 --prompt 1000"
