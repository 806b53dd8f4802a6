#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised
times: compare SHARES, not absolutes).   python tools/launch_summary.py launches.csv"""
import csv
import sys
from collections import defaultdict

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
d = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    n = r["Kernel Name"].split("(")[0][-70:]
    d[n][0] += 1
    d[n][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in d.values())
print(f"{'launches':>8s} {'total ms':>10s} {'share':>6s}  kernel")
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[0]:8d} {v[1]:10.3f} {100 * v[1] / tot:5.1f}%  {k}")
