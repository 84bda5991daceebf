#!/usr/bin/env python
"""Minimal driver for ncu captures of the MultiScaleRoIAlign kernels (bench.py shapes: B=8,
608x1024 padded image, 512 RoIs/img, C=256)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dgod_b200 import ops, synth

nhwc = "--nchw" not in sys.argv
dtype = torch.bfloat16 if "--bf16" in sys.argv else torch.float32
DEV = torch.device("cuda")
B, C, H, W, per = 8, 256, 608, 1024, 512
feats = [f.to(DEV) for f in synth.random_features(B, C, H, W, 0, dtype=dtype)]
if nhwc:
    feats = [f.contiguous(memory_format=torch.channels_last) for f in feats]
feats = [f.requires_grad_(True) for f in feats]
boxes = [synth.random_boxes(per, H, W, synth.gen(10 + i)) for i in range(B)]
rois = synth.rois_from_boxes(boxes).to(DEV)
offs = ops._offsets([per] * B, DEV)
for _ in range(3):
    out = ops.multiscale_roi_align(feats, rois, [1 / 4, 1 / 8, 1 / 16, 1 / 32], 7, 2, 2, 5, roi_img_offsets=offs)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
print("ok")
