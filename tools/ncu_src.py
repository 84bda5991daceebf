#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top source lines by stall samples / instructions.
usage: ncu -i rep --page source --csv --kernel-name regex:X > f.csv ; python tools/ncu_src.py f.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
src_i, samp_i, inst_i = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
tot = 0
items = []
for r in rows[h + 1:]:
    if len(r) <= samp_i or not r[samp_i]:
        continue
    try:
        s = int(r[samp_i]); ins = int(r[inst_i])
    except ValueError:
        continue
    tot += s
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    items.append((s, ins, r[src_i][:90], st))
items.sort(reverse=True)
print("total samples", tot, "instr", sum(i[1] for i in items))
for s, ins, src, st in items[:n]:
    print(f"{100*s/max(tot,1):5.1f}% inst={ins:9d} {src:90s} {st}")
