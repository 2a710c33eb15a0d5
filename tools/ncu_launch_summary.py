#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: usage: ncu_launch_summary.py file.csv"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    tot[name][0] += 1
    tot[name][1] += v
T = sum(v[1] for v in tot.values())
print("%d launches, %.1f us total" % (sum(v[0] for v in tot.values()), T))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-72s n=%4d total=%9.1fus avg=%7.2fus share=%.3f" % (k[:72], v[0], v[1], v[1] / v[0], v[1] / T))
