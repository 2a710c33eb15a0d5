#!/bin/bash
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_llama_gpu.py tests/test_layers_gpu.py -m gpu -q -x --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -4
timeout 300 python bench.py --model llama-3.2-1b --steps 64 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('1B bf16', round(d['value'],1),'tok/s', round(d['ms_per_step'],4),'ms roof', round(d['roofline']['frac'],3), 'step_roof', round(d['step_roofline']['frac'],3), d.get('latency'))"
timeout 300 python bench.py --steps 64 --warmup 4 --no-cpu-baseline --requests 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('8B gptq', round(d['value'],1),'tok/s', round(d['ms_per_step'],4),'ms')"
