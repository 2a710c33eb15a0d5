#!/usr/bin/env python
"""Print the key metrics of an .ncu-rep (raw page) per kernel launch: usage: ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
for k in KEYS:
    if k in idx:
        print(k[:72].ljust(74), [r[idx[k]][:18] for r in data])
for h in hdr:
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        vals = [r[idx[h]][:6] for r in data]
        if any(float(v or 0) > 0.15 for v in vals):
            print(h.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', '').ljust(74), vals)
