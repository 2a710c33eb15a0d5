#!/bin/bash
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
echo "== TS kernel timing ablations (ZL_TC_DBG: 1 no x loads, 2 no dequant, 4 no MMA, 8 no weight loads)"
for shape in "28672 4096 32" "28672 4096 128" "6144 4096 32"; do
for dbg in 0 1 2 4 8 3 6 7 15; do
ZL_TC_DBG=$dbg timeout 120 python tools/tc_bench.py --one $shape 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('shape $shape dbg $dbg', round(d['us'],1),'us', round(d['hbm_frac'],3))
"
done
done
echo "== splits forced (28672x4096 M=32)"
for s in 1 2 3 4; do ZL_TC_SPLITS=$s timeout 120 python tools/tc_bench.py --one 28672 4096 32 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('splits $s', round(d['us'],1),'us', round(d['hbm_frac'],3))
"; done
echo "== timeline mask 70"; ZL_NO_PDL_MASK=70 timeout 300 python tools/trace_step.py > $O/timeline70.txt 2>&1; tail -4 $O/timeline70.txt
echo "== timeline mask 66 (o with PDL)"; ZL_NO_PDL_MASK=66 timeout 300 python tools/trace_step.py > $O/timeline66.txt 2>&1; tail -4 $O/timeline66.txt
echo "== ncu TS kernel M=32 28672x4096"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_w4a16_ts -s 3 -c 1 -o $O/ts_m32 -f python tools/tc_bench.py --one 28672 4096 32 > $O/ncu_ts.log 2>&1; tail -2 $O/ncu_ts.log; ls -la $O/*.ncu-rep
