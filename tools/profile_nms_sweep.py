#!/usr/bin/env python
"""Driver for an ncu launch list of batched_nms at the op-sweep sizes (BASELINE configs[3]):
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file f.csv \
        python tools/profile_nms_sweep.py 100000"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dgod_b200 import ops, synth

DEV = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
g = synth.gen(n)
boxes = synth.random_boxes(n, 800, 1333, g).to(DEV)
scores = synth.distinct_scores(n, g).to(DEV)
idxs = torch.randint(0, 5, (n,), generator=g).to(DEV)
for it in range(3):
    if it == 2:
        torch.cuda.profiler.start()
    ops.nms_segments(boxes, scores, idxs, [n], 0.7)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
