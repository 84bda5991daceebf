#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu -i rep --page source --csv --print-source cuda,sass > f.csv ; python tools/ncu_lines.py f.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
agg = collections.OrderedDict()
cur = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        samp_i, inst_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr is None or len(r) <= samp_i:
        continue
    if r[0]:
        cur = (r[0], r[1].strip()[:100])
        agg.setdefault(cur, [0, 0, collections.Counter()])
    if cur is None or not r[2]:
        continue
    try:
        s, ins = int(r[samp_i] or 0), int(r[inst_i] or 0)
    except ValueError:
        continue
    a = agg[cur]
    a[0] += s; a[1] += ins
    for i in stall:
        v = int(r[i] or 0)
        if v: a[2][hdr[i][6:]] += v
tot = sum(a[0] for a in agg.values())
print("total samples", tot, "warp instructions", sum(a[1] for a in agg.values()))
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    top = ", ".join(f"{k} {v}" for k, v in a[2].most_common(3))
    print(f"{100 * a[0] / max(tot, 1):5.1f}% inst={a[1]:9d} L{ln:>4s} {src:100s} [{top}]")
