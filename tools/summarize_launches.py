#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "second": 1e9}.get(row["Metric Unit"], 1)
    agg[row["Kernel Name"]][0] += 1
    agg[row["Kernel Name"]][1] += v
    tot += v
print(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.2f} ms of kernel time (cold-cache, serialised)")
print("share%  total_ms  launches  avg_us  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * t / tot:6.2f} {t / 1e6:9.2f} {n:9d} {t / n / 1e3:8.1f}  {k[:110]}")
