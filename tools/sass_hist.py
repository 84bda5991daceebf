#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libdgod_b200.so (cuobjdump -sass): the evidence for which hardware paths a kernel
uses (UTMALDG = tensor-map TMA load, UBLKCP = 1-D bulk copy, UBLKRED = bulk reduce, SYNCS = mbarrier, FFMA2 = packed fp32,
UCGABAR_* = cluster barrier, REDUX / VOTE / MATCH = warp collectives, ATOMS / ATOMG / RED = atomics).

    python tools/sass_hist.py [--top 12] > profiles/r02_sass_histogram.txt"""
import argparse
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEY = ["UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "FFMA2", "FFMA", "HMMA", "REDUX", "VOTE",
       "MATCH", "SHFL", "ATOMS", "ATOMG", "RED", "LDS", "STS", "LDG", "STG", "BRX"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=10)
    a = ap.parse_args()
    out = subprocess.run(["cuobjdump", "-sass", str(ROOT / "dgod_b200" / "libdgod_b200.so")], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    print("# SASS opcode histogram per kernel of dgod_b200/libdgod_b200.so (sm_100a), `python tools/sass_hist.py`")
    print("# columns: total instructions | the mnemonics that prove a hardware path | the most frequent other opcodes\n")
    for name, c in kernels.items():
        keyed = " ".join(f"{k}:{c[k]}" for k in KEY if c[k])
        rest = " ".join(f"{k}:{v}" for k, v in c.most_common() if k not in KEY)[: 16 * a.top]
        print(f"{name}\n    total {sum(c.values())} | {keyed or '-'} | {rest}")


if __name__ == "__main__":
    sys.exit(main())
