#!/usr/bin/env python
"""Minimal driver for ncu launch lists of the two NMS pipelines at the training step's shapes:
rpn_proposals (B=8, 608x1024) and the box-head post-processing NMS (8 images x 512 RoIs x 8 classes)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dgod_b200 import ops, synth
from dgod_b200.detector import make_cell_anchors

DEV = torch.device("cuda")
cells = [c.tolist() for c in make_cell_anchors(((32,), (64,), (128,), (256,), (512,)), ((0.5, 1.0, 2.0),) * 5)]
h, w, B = 608, 1024, 8
g = synth.gen(3)
grids = [(-(-h // s), -(-w // s)) for s in (4, 8, 16, 32, 64)]
strides = [(h // gh, w // gw) for gh, gw in grids]
obj = [torch.randn(B, 3, gh, gw, generator=g).to(DEV) for gh, gw in grids]
dl = [(torch.randn(B, 12, gh, gw, generator=g) * 0.2).to(DEV) for gh, gw in grids]
sizes = torch.tensor([[600.0, 999.0]] * B, device=DEV)
# box head: 512 proposals x 8 foreground classes per image, scores like a random-init softmax (~1/9)
n = B * 512 * 8
boxes = synth.random_boxes(n, 600, 999, g).to(DEV)
scores = (torch.rand(n, generator=g) * 0.2).to(DEV)
labels = (torch.arange(n) % 8 + 1).to(DEV)
valid = (scores > 0.05).to(torch.uint8)
if "--time" in sys.argv:      # CUDA-event times (L2 flushed between iterations, median of 20), no profiler
    import statistics
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(20):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return statistics.median(ts), min(ts)

    print("rpn_proposals  B8 608x1024            us median/min: %.1f / %.1f" % timed(lambda: ops.rpn_proposals(obj, dl, sizes, strides, cells, 2000, 2000, 0.7)))
    print("nms_segments   8 x 4096, 8 classes    us median/min: %.1f / %.1f" % timed(lambda: ops.nms_segments(boxes, scores, labels, [512 * 8] * B, 0.5, valid=valid, max_out_per_seg=100)))
    for n in (10000, 100000):
        gg = synth.gen(n)
        bx = synth.random_boxes(n, 800, 1333, gg).to(DEV)
        sc = synth.distinct_scores(n, gg).to(DEV)
        ix = torch.randint(0, 5, (n,), generator=gg).to(DEV)
        print("batched_nms    %6d boxes, 5 groups    us median/min: %.1f / %.1f" % ((n,) + timed(lambda: ops.nms_segments(bx, sc, ix, [n], 0.7))))
    sys.exit(0)
for it in range(3):
    torch.cuda.profiler.start() if it == 2 else None
    ops.rpn_proposals(obj, dl, sizes, strides, cells, 2000, 2000, 0.7)
    torch.cuda.synchronize()
    ops.nms_segments(boxes, scores, labels, [512 * 8] * B, 0.5, valid=valid, max_out_per_seg=100)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
