#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time, launches and
share per kernel, dgod_b200 kernels marked.
usage: python tools/ncu_launches.py launches.csv [top_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for r in rows:
    if len(r) != len(hdr) or r is hdr or not r[0].isdigit():
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    a = agg[name]
    a[0] += 1
    a[1] += v * scale
    n += 1
tot = sum(a[1] for a in agg.values())
ours = sum(a[1] for k, a in agg.items() if k.startswith("dgod::"))
print(f"{n} launches, {tot / 1e3:.2f} ms of kernel time (serialised, cold-cache); dgod_b200 kernels: "
      f"{sum(a[0] for k, a in agg.items() if k.startswith('dgod::'))} launches, {ours / 1e3:.2f} ms = {100 * ours / tot:.1f} %")
print(f"{'kernel':78s} {'n':>6s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[:78]:78s} {a[0]:6d} {a[1]:10.1f} {a[1] / a[0]:9.1f} {100 * a[1] / tot:5.1f}%")
