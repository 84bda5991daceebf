#!/usr/bin/env python
"""A/B of the MultiScaleRoIAlign backward algorithms through the C ABI (bench.py shapes by default).

    python tools/bwd_check.py [--per 512] [--batch 8] [--algos 3,4] [--dtype f32|bf16] [--hw 608x1024] [--iters 15]

Prints kernel_us (CUDA events around the C-ABI call, L2 flushed between iterations, median) per
algorithm, the max relative difference of every algorithm's gradients against the first one, and
whether two runs of each algorithm are bitwise equal (determinism)."""
import argparse
import ctypes as C
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dgod_b200 import _lib, ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--per", type=int, default=512)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--algos", default="3,4")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--hw", default="608x1024")
ap.add_argument("--iters", type=int, default=15)
ap.add_argument("--offsets", type=int, default=1)
ap.add_argument("--lib", default="", help="a variant library built by tools/roi_variants.py (tools/_variants/lib_<tag>.so)")
ap.add_argument("--warps", type=int, default=28, help="consumer warps per CTA of the timing build")
ap.add_argument("--timing", action="store_true", help="read the per-CTA cycle counters of a -DDGOD_OWN_TIMING build")
a = ap.parse_args()

if a.lib:
    _lib.LIB_PATH = Path(a.lib).resolve()
DEV = torch.device("cuda")
dtype = torch.float32 if a.dtype == "f32" else torch.bfloat16
H, W = (int(v) for v in a.hw.split("x"))
B_, Cc = a.batch, 256
feats = [f.to(DEV).contiguous(memory_format=torch.channels_last) for f in synth.random_features(B_, Cc, H, W, 0, dtype=dtype)]
boxes = [synth.random_boxes(a.per, H, W, synth.gen(10 + i)) for i in range(B_)]
rois = synth.rois_from_boxes(boxes).to(DEV)
offs = ops._offsets([a.per] * B_, DEV) if a.offsets else None
lib = _lib.load()
K = rois.shape[0]
go = torch.randn(K, Cc, 7, 7, generator=synth.gen(99)).to(DEV, dtype)
cfg, _ = ops._roi_config(feats, [1 / 4, 1 / 8, 1 / 16, 1 / 32], 7, 7, 2, False, 2, 5, 224.0, 4.0)
cfg.channels_last = 1
wsb = lib.dgod_msroi_align_bwd_workspace_bytes_cfg(C.byref(cfg), K)
ws = torch.zeros(wsb, dtype=torch.uint8, device=DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
st = ops._stream()
alg_bytes = K * Cc * 49 * go.element_size() + 20 * K + sum(f.numel() for f in feats) * go.element_size()
print(f"K={K} dtype={a.dtype} {H}x{W} alg_bytes={alg_bytes / 1e6:.1f} MB workspace={wsb / 1e6:.1f} MB")

base = None
for algo in (int(v) for v in a.algos.split(",")):
    grads = [torch.full_like(f, float("nan")) for f in feats]
    gptrs, keep = ops._level_ptrs(grads)

    def bwd():
        ops.check(lib.dgod_msroi_align_bwd(C.byref(cfg), ops._p(go), ops._p(rois), K, ops._p(offs), gptrs, algo, ops._p(ws), wsb, st))

    for _ in range(3):
        bwd()
    torch.cuda.synchronize()
    first = [g.clone() for g in grads]
    ts = []
    for _ in range(a.iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bwd(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = statistics.median(ts)
    det = all(torch.equal(x, y) for x, y in zip(first, grads))
    msg = f"algo {algo}: {us:8.1f} us  {alg_bytes / 1e3 / us:7.1f} GB/s  min {min(ts):.1f}  deterministic={det}"
    res = [g.float() for g in grads]
    if base is None:
        base = res
    else:
        errs = [((x - y).abs().max() / y.abs().max()).item() for x, y in zip(res, base)]
        msg += "  relerr vs first: " + " ".join(f"{e:.2e}" for e in errs)
    print(msg, flush=True)
    if a.timing and algo == 4:
        t = ws[1024:1024 + 148 * a.warps * 4 * 8].view(torch.int64).view(148, a.warps, 4).cpu().double()
        tot, wait, store, pairs = (t[..., i] for i in range(4))
        busy = tot - wait - store
        print(f"  per CTA (slowest warp) total cycles min/mean/max {tot.max(1).values.min():.0f}/{tot.max(1).values.mean():.0f}/{tot.max(1).values.max():.0f}  "
              f"pairs per CTA min/mean/max {pairs[:, 0].min():.0f}/{pairs[:, 0].mean():.0f}/{pairs[:, 0].max():.0f}")
        print("  per warp position (mean over CTAs): busy " + " ".join(f"{v / 1e3:.0f}k" for v in busy.mean(0)))
        print("                                      wait " + " ".join(f"{v / 1e3:.0f}k" for v in wait.mean(0)))
        print(f"  store mean {store.mean():.0f}; busiest warp of a CTA / mean warp of that CTA = {(busy.max(1).values / busy.mean(1)).mean():.2f}")
