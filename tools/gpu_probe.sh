#!/bin/bash
mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print('$1', {k:round(d[k],4) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','kernels_per_step')}, 'gemm us/launch', round(d['roofline']['us_per_launch'],2), 'frac', round(d['roofline']['frac'],3))
"; }
echo "== stream probe (no dequant/MMA)"; ZL_W4_DEBUG=1 run probe ""
ZL_W4_DEBUG=1 ZL_DEBUG_SKIP=1 run probe_noattn ""
echo "== ncu v2"
TAG=r01b
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_w4a16_v2 -s 300 -c 4 -f -o gpurun_out/${TAG}_prof_w4a16 python bench.py --steps 2 --warmup 3 --prompt 2 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_attn -s 40 -c 2 -f -o gpurun_out/${TAG}_prof_attn python bench.py --steps 2 --warmup 3 --prompt 200 --no-cpu-baseline > gpurun_out/${TAG}_ncu_attn.log 2>&1; echo "attn rc=$?"
