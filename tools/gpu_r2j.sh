#!/bin/bash
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
echo "== trace 28672x4096 M=32"; timeout 200 python tools/tc_trace.py 28672 4096 32 > $O/trace_big32.txt 2>&1; head -45 $O/trace_big32.txt
echo "== trace, everything ablated (dbg 15)"; ZL_TC_DBG=15 timeout 200 python tools/tc_trace.py 28672 4096 32 > $O/trace_big32_empty.txt 2>&1; head -45 $O/trace_big32_empty.txt
echo "== trace 28672x4096 M=128"; timeout 200 python tools/tc_trace.py 28672 4096 128 > $O/trace_big128.txt 2>&1; sed -n 1,30p $O/trace_big128.txt
