#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: python tools/summarize_launches.py profiles/r1_launches_bench_one_pair.csv"""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.DictReader(rows)
agg = collections.OrderedDict()
k1_l0 = []
total = 0.0
n = 0
for r in rd:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("ugsm::", "").replace("void ", "")
    us = float(r["Metric Value"]) / 1e3
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] = max(a[2], us)
    total += us
    n += 1
    if name.startswith("k1_correlate<") and r["Grid Size"].replace(" ", "") == "(83,204,1)":
        k1_l0.append(us)
print("%d launches, sum of kernel durations %.2f ms (cold-cache, serialised: compare shares)\n" % (n, total / 1e3))
for name, (cnt, s, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s n=%4d  sum=%9.1f us  share=%5.1f%%  max=%8.1f us" % (name, cnt, s, 100 * s / total, mx))
if k1_l0:
    print("\nlevel-0 k1_correlate launches (grid 83x204): %s us -> share of the pair %.1f%%" % (
        [round(v, 1) for v in k1_l0], 100 * sum(k1_l0) / total))
