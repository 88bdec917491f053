#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
n_steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f"{'kernel':64s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:64]:64s} {len(v):8d} {sum(v):10.3f} {sum(v) / len(v) * 1e3:10.1f} {sum(v) / tot * 100:6.1f}%")
print(f"{'TOTAL':64s} {sum(len(v) for v in agg.values()):8d} {tot:10.3f}   (over {n_steps:g} captured steps: {tot / n_steps:.3f} ms/step)")
