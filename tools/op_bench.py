#!/usr/bin/env python
"""Op sweep of BASELINE.json configs[3]: per-kernel time and achieved GB/s against the roofline.

    python tools/op_bench.py [--ops roi,nms,match,fcos,rpn,grl] [--iters 20] [--tv] [--json out.json]

Timing: CUDA events on the launching stream, 3 warm-up launches, an L2 flush (a 256 MB write)
between timed iterations, median over --iters.  Algorithmic bytes follow SURVEY.md §8d.  --tv also
times the stock torchvision CUDA op on the same inputs (the "existing sm_100 recompile" bar).
"""
from __future__ import annotations

import argparse
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from dgod_b200 import ops, synth  # noqa: E402
from dgod_b200.detector import make_cell_anchors  # noqa: E402

DEV = torch.device("cuda")
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    _flush.fill_(1)


LAST_KERNEL_US = None   # median time between events placed directly around the C-ABI call(s) of the last time_op


def time_op(fn, iters, warmup=3, flush=True):
    """Median op time in us (events around the Python call, so it includes the wrapper's small torch
    kernels and host gaps) — and, in LAST_KERNEL_US, the time of the C-ABI call alone."""
    global LAST_KERNEL_US
    for _ in range(warmup):
        fn()
    ts, ks = [], []
    ops.KernelTimer.enabled = True
    for _ in range(iters):
        if flush:
            flush_l2()
        ops.KernelTimer.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        k = sum(v[1] for v in ops.KernelTimer.summary().values())
        if k > 0:
            ks.append(k)
    ops.KernelTimer.enabled = False
    LAST_KERNEL_US = statistics.median(ks) * 1e3 if ks else None
    return statistics.median(ts) * 1e3  # us


def peak():
    f = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(f.read_text())["hbm_gbs"]) if f.exists() else 6650.0


def row(name, size, us, nbytes, extra=None):
    """GB/s and roofline fraction use the kernel time (C-ABI call) when it was captured."""
    k_us = LAST_KERNEL_US if (LAST_KERNEL_US and not name.startswith("torchvision")) else us
    gbs = nbytes / 1e3 / k_us
    r = {"op": name, "size": size, "op_us": round(us, 2), "kernel_us": round(k_us, 2), "alg_MB": round(nbytes / 1e6, 2),
         "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak(), 4)}
    if extra:
        r.update(extra)
    print(json.dumps(r), flush=True)
    return r


def bench_roi(args, out):
    B, C, H, W = args.batch, 256, args.height, args.width
    for dtype in (torch.float32, torch.bfloat16):
        feats = [f.to(DEV) for f in synth.random_features(B, C, H, W, 0, dtype=dtype)]
        esz = feats[0].element_size()
        fbytes = sum(f.numel() for f in feats) * esz
        scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
        for per in args.rois:
            boxes = [synth.random_boxes(per, H, W, synth.gen(10 + i)) for i in range(B)]
            rois = synth.rois_from_boxes(boxes).to(DEV)
            offs = ops._offsets([per] * B, DEV)
            K = rois.shape[0]
            for nhwc in (False, True):
                fs = [f.contiguous(memory_format=torch.channels_last) if nhwc else f for f in feats]
                fs = [f.detach().requires_grad_(True) for f in fs]
                fwd = lambda: ops.multiscale_roi_align(fs, rois, scales, 7, 2, 2, 5, roi_img_offsets=offs)
                with torch.no_grad():
                    us = time_op(fwd, args.iters)
                nb = ops._roi_bytes([f.numel() for f in fs], esz, K, C, 7, 7, 2)
                tag = f"{'bf16' if dtype == torch.bfloat16 else 'f32'}-{'nhwc' if nhwc else 'nchw'}"
                out.append(row("msroi_align_fwd", f"B{B} {per}/img {tag}", us, nb))
                if dtype == torch.float32 or nhwc or True:
                    o = fwd()
                    go = torch.randn_like(o)
                    def bwd():
                        for f in fs:
                            f.grad = None
                        o.backward(go, retain_graph=True)
                    try:
                        us = time_op(bwd, args.iters)
                        nbb = K * C * 49 * esz + 20 * K + fbytes
                        out.append(row("msroi_align_bwd", f"B{B} {per}/img {tag}", us, nbb))
                    except RuntimeError as e:
                        print(json.dumps({"op": "msroi_align_bwd", "size": tag, "error": str(e)[:100]}))
            if args.tv and dtype == torch.float32:
                import torchvision
                pool = torchvision.ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2)
                x = {str(i): f.detach().requires_grad_(True) for i, f in enumerate(feats)}
                bl = [b.to(DEV) for b in boxes]
                fwd_tv = lambda: pool(x, bl, [(H, W)] * B)
                with torch.no_grad():
                    us = time_op(fwd_tv, args.iters)
                out.append(row("torchvision_msroi_fwd", f"B{B} {per}/img f32-nchw", us, nb))
                o = fwd_tv()
                go = torch.randn_like(o)
                def bwd_tv():
                    for f in x.values():
                        f.grad = None
                    o.backward(go, retain_graph=True)
                us = time_op(bwd_tv, args.iters)
                out.append(row("torchvision_msroi_bwd", f"B{B} {per}/img f32-nchw", us, K * C * 49 * 4 + 20 * K + fbytes))


FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e-3      # CUDA-core fp32 peak of a B200


def pair_stats(run_sizes):
    """NMS is O(n^2) pair tests over O(n) bytes: pair evaluations per second and, at ~14 flop per
    pair, the fraction of the fp32 CUDA-core peak, from the kernel time of the row being built."""
    pairs = sum(r * (r - 1) // 2 for r in run_sizes)
    k_us = LAST_KERNEL_US or 1.0
    return {"pairs": pairs, "Gpairs/s": round(pairs / 1e3 / k_us, 2),
            "frac_of_fp32_peak": round(pairs * 14 / 1e6 / k_us / FP32_PEAK_TFLOPS, 4)}


def bench_nms(args, out):
    for n in args.nms:
        g = synth.gen(n)
        boxes = synth.random_boxes(n, 800, 1333, g).to(DEV)
        scores = synth.distinct_scores(n, g).to(DEV)
        idxs = torch.randint(0, 5, (n,), generator=g).to(DEV)
        keep = ops.batched_nms(boxes, scores, idxs, 0.7)
        nb = 28 * n + 8 * keep.numel()
        us = time_op(lambda: ops.nms_segments(boxes, scores, idxs, [n], 0.7), args.iters)
        out.append(row("batched_nms", f"{n} boxes, 5 groups, thr 0.7", us, nb,
                       {"kept": keep.numel(), **pair_stats(torch.bincount(idxs.cpu()).tolist())}))
        if args.tv:
            import torchvision
            us = time_op(lambda: torchvision.ops.batched_nms(boxes, scores, idxs, 0.7), args.iters)
            out.append(row("torchvision_batched_nms", f"{n} boxes, 5 groups, thr 0.7", us, nb))
    # the RPN shape: 8 images x 8304 candidates in 5 level groups, one call
    counts = [8304] * 8
    n = sum(counts)
    g = synth.gen(1)
    boxes = synth.random_boxes(n, 608, 1024, g).to(DEV)
    scores = synth.distinct_scores(n, g).to(DEV)
    idxs = torch.randint(0, 5, (n,), generator=g).to(DEV)
    us = time_op(lambda: ops.nms_segments(boxes, scores, idxs, counts, 0.7, max_out_per_seg=2000), args.iters)
    runs = [c for i in range(8) for c in torch.bincount(idxs[i * 8304:(i + 1) * 8304].cpu(), minlength=5).tolist()]
    out.append(row("batched_nms", "8 images x 8304 boxes, 5 groups (one call)", us, 28 * n + 8 * 8 * 2000, pair_stats(runs)))


def bench_match(args, out):
    for (b, m, n, hi, lo, lq, tag) in [(8, 20, 155520, 0.7, 0.3, True, "RPN anchors 608x1024"),
                                      (8, 20, 268569, 0.7, 0.3, True, "RPN anchors 800x1344"),
                                      (8, 100, 2000, 0.5, 0.5, False, "RoI heads 100 GT x 2000")]:
        gts = [synth.random_boxes(m, 608, 1024, synth.gen(i)).to(DEV) for i in range(b)]
        if n > 10000:
            boxes = synth.random_boxes(n, 608, 1024, synth.gen(99)).to(DEV)
            fn = lambda: ops.match_boxes(gts, boxes, hi, lo, lq, want=("labels_f32", "matched_boxes"))
            nb = b * (16 * m + 8 * n + 4 * n + 16 * n) + 16 * n
        else:
            boxes = [synth.random_boxes(n, 608, 1024, synth.gen(50 + i)).to(DEV) for i in range(b)]
            labels = [torch.randint(1, 9, (m,), generator=synth.gen(i)).to(DEV) for i in range(b)]
            fn = lambda: ops.match_boxes(gts, boxes, hi, lo, lq, gt_labels=labels, want=("labels_i64", "clamped_idx"))
            nb = b * (16 * m + 16 * n + 24 * n)
        out.append(row("iou_match", f"B{b} {m} GT x {n} ({tag})", time_op(fn, args.iters), nb))


def bench_fcos(args, out):
    from dgod_b200.detector import grid_anchors
    for (h, w) in [(800, 1344), (608, 1024)]:
        strides = (8, 16, 32, 64, 128)
        grids = [(-(-h // s), -(-w // s)) for s in strides]
        cells = [torch.tensor([[-4.0 * s, -4.0 * s, 4.0 * s, 4.0 * s]]) for s in strides]   # fcos.py:467-468
        npl = [gh * gw for gh, gw in grids]
        a = grid_anchors(cells, grids, [(s, s) for s in strides], DEV).float()
        gts = [synth.random_boxes(20, h, w, synth.gen(i)).to(DEV) for i in range(8)]
        labels = [torch.randint(1, 9, (20,), generator=synth.gen(i)).to(DEV) for i in range(8)]
        n = a.shape[0]
        us = time_op(lambda: ops.fcos_assign(a, gts, npl, 1.5), args.iters)
        out.append(row("fcos_assign", f"B8 {n} locations x 20 GT", us, 8 * (16 * n + 16 * 20 + 8 * n)))
        us = time_op(lambda: ops.fcos_assign(a, gts, npl, 1.5, gt_labels=labels, num_classes=9), args.iters)
        out.append(row("fcos_assign+targets", f"B8 {n} locations x 20 GT", us, 8 * (16 * n + 16 * 20 + 8 * n + 8 * n + 16 * n + 36 * n)))
        # loss tail (fcos.py:149-202): fused kernels vs the reference's ATen chain on the same device
        from dgod_b200.dg_fcos import FCOSHead
        g = synth.gen(77)
        ho = {"cls_logits": (torch.randn(8, n, 9, generator=g) * 2).to(DEV).requires_grad_(True),
              "bbox_regression": (torch.rand(8, n, 4, generator=g) * 2 + 0.05).to(DEV).requires_grad_(True),
              "bbox_ctrness": torch.randn(8, n, 1, generator=g).to(DEV).requires_grad_(True)}
        assigned = ops.fcos_assign(a, gts, npl, 1.5, gt_labels=labels, num_classes=9)
        head = FCOSHead(256, 1, 9).to(DEV)
        alg = 8 * n * (4 * 14 + 8 + 16) + 16 * n

        def step(fused):
            head.fused_loss = fused
            for t in ho.values():
                t.grad = None
            d = head.compute_loss(None, ho, [a] * 8, assigned)
            (d["classification"] + d["bbox_regression"] + d["bbox_ctrness"]).backward()

        us = time_op(lambda: step(True), args.iters)
        out.append(row("fcos_loss fwd+bwd (fused; through head.compute_loss, loss sum and the autograd engine)", f"B8 {n} locations", us, 3 * alg))
        # the op calls alone (what the wrapper adds to the kernels: output allocation, one ctypes call)
        with torch.no_grad():
            dt = [ho["cls_logits"].detach(), ho["bbox_regression"].detach(), ho["bbox_ctrness"].detach()]
            _, cls_t, box_t, _ = assigned
            us = time_op(lambda: ops.fcos_loss(dt[0], dt[1], dt[2], a, cls_t, box_t), args.iters)
            out.append(row("fcos_loss forward (op call)", f"B8 {n} locations", us, alg))
            losses = ops.fcos_loss(dt[0], dt[1], dt[2], a, cls_t, box_t)
            gl = torch.ones(3, device=DEV)
            us = time_op(lambda: ops._fcos_loss_bwd_op._init_fn(dt[0], dt[1], dt[2], a.float().contiguous(), cls_t.contiguous(),
                                                                box_t.float().contiguous(), 0.25, losses, gl), args.iters)
            out.append(row("fcos_loss backward (op call)", f"B8 {n} locations", us, 2 * alg))
        if args.tv:
            t = time_op(lambda: step(False), args.iters)
            out.append(row("torchvision_ops_fcos_loss fwd+bwd (ATen chain of fcos.py:149-202)", f"B8 {n} locations", t, 3 * alg))


def bench_transform(args, out):
    """GeneralizedRCNNTransform (TV transform.py:102-153) on 8 x 3x800x1333 -> 600x999 -> 608x1024."""
    from dgod_b200.detector import FusedTransform
    imgs = [i.to(DEV) for i in synth.random_images(8, 800, 1333, 0)]
    nbytes = 8 * 3 * 800 * 1333 * 4 + 8 * 3 * 608 * 1024 * 4
    us = time_op(lambda: ops.image_batch(imgs, [0.0] * 3, [1.0] * 3, 600, 1200), args.iters)
    out.append(row("image_batch (normalize+resize+pad)", "8 x 3x800x1333 -> 608x1024", us, nbytes))
    if args.tv:
        from torchvision.models.detection.transform import GeneralizedRCNNTransform
        tv = GeneralizedRCNNTransform(600, 1200, [0.0] * 3, [1.0] * 3).to(DEV)
        us = time_op(lambda: tv(imgs), args.iters)
        out.append(row("torchvision_transform", "8 x 3x800x1333 -> 608x1024", us, nbytes))


def bench_rpn(args, out):
    cells = [c.tolist() for c in make_cell_anchors(((32,), (64,), (128,), (256,), (512,)), ((0.5, 1.0, 2.0),) * 5)]
    for (h, w) in [(608, 1024), (800, 1344)]:
        B = 8
        g = synth.gen(3)
        grids = [(-(-h // s), -(-w // s)) for s in (4, 8, 16, 32, 64)]
        strides = [(h // gh, w // gw) for gh, gw in grids]
        obj = [torch.randn(B, 3, gh, gw, generator=g).to(DEV) for gh, gw in grids]
        dl = [(torch.randn(B, 12, gh, gw, generator=g) * 0.2).to(DEV) for gh, gw in grids]
        sizes = torch.tensor([[h, w]] * B, dtype=torch.float32, device=DEV)
        A = sum(3 * gh * gw for gh, gw in grids)
        k = sum(min(2000, 3 * gh * gw) for gh, gw in grids)
        us = time_op(lambda: ops.rpn_proposals(obj, dl, sizes, strides, cells, 2000, 2000, 0.7), args.iters)
        out.append(row("rpn_proposals", f"B8 A={A} k={k}", us, B * (4 * A + k * 16 + 2000 * 20)))


def bench_grl(args, out):
    for shape in [(4096, 1024), (8, 256, 152, 256)]:
        x = torch.randn(*shape, device=DEV)
        us = time_op(lambda: ops._grl_scale_op(x, 0.1), args.iters)
        out.append(row("grl_scale", "x".join(map(str, shape)), us, 2 * x.numel() * 4))


def bench_sampler(args, out):
    """BalancedPositiveNegativeSampler (TV _utils.py:11-71): the one-launch kernel vs the two-topk torch formulation."""
    for (n, P, S, dt, tag) in [(155520, 128, 256, torch.float32, "RPN anchors 608x1024"), (268569, 128, 256, torch.float32, "RPN anchors 800x1344"),
                               (2020, 128, 512, torch.int64, "RoI heads 2000 proposals + 20 GT")]:
        g = synth.gen(n)
        r = torch.rand(8, n, generator=g)
        lab = torch.full((8, n), -1.0)
        lab[r < 0.6] = 0.0
        lab[r < 0.002] = 1.0
        labels, keys = lab.to(dt).to(DEV), torch.rand(8, n, generator=g).to(DEV)
        us = time_op(lambda: ops.balanced_sample(labels, keys, P, S), args.iters)
        out.append(row("balanced_sample", f"B8 x {n} ({tag})", us, 8 * n * (labels.element_size() + 4) + 8 * (P + S) * 9))
        if args.tv:
            def two_topk():
                pk = torch.where(labels >= 1, keys, torch.full_like(keys, 2.0))
                nk = torch.where(labels == 0, keys, torch.full_like(keys, 2.0))
                return torch.topk(pk, min(P, n), dim=1, largest=False), torch.topk(nk, min(S, n), dim=1, largest=False)
            us = time_op(two_topk, args.iters)
            out.append(row("torch_two_topk_sampler", f"B8 x {n} ({tag})", us, 8 * n * (labels.element_size() + 4)))


def bench_fcos_post(args, out):
    """FCOS eval candidates (fcos.py:576-597): one launch vs the per-image, per-level ATen chain of the reference."""
    from dgod_b200.detector import grid_anchors
    for (h, w) in [(608, 1024), (800, 1344)]:
        strides = (8, 16, 32, 64, 128)
        grids = [(-(-h // s), -(-w // s)) for s in strides]
        cells = [torch.tensor([[-4.0 * s, -4.0 * s, 4.0 * s, 4.0 * s]]) for s in strides]
        npl = [gh * gw for gh, gw in grids]
        a = grid_anchors(cells, grids, [(s, s) for s in strides], DEV).float()
        n = a.shape[0]
        g = synth.gen(5)
        cl = (torch.randn(8, n, 9, generator=g) * 2.5 - 2.0).to(DEV)
        rg = (torch.rand(8, n, 4, generator=g) * 3 + 0.1).to(DEV)
        ct = (torch.randn(8, n, 1, generator=g) * 1.5).to(DEV)
        sizes = torch.tensor([[float(h), float(w)]] * 8, device=DEV)
        us = time_op(lambda: ops.fcos_candidates(cl, rg, ct, a, npl, sizes, 0.2, 1000), args.iters)
        out.append(row("fcos_candidates", f"B8 {n} locations x 9 classes, top-1000 x 5 levels", us, 8 * n * 40 + 8 * 5000 * 45))
        if args.tv:
            def chain():
                res = []
                for i in range(8):
                    off = 0
                    for nl in npl:
                        s_ = torch.sqrt(torch.sigmoid(cl[i, off:off + nl]) * torch.sigmoid(ct[i, off:off + nl])).flatten()
                        keep = s_ > 0.2
                        s2 = s_[keep]
                        idx = torch.where(keep)[0]
                        k = min(1000, idx.numel())
                        s2, j = s2.topk(k)
                        idx = idx[j]
                        ai = torch.div(idx, 9, rounding_mode="floor")
                        an, r = a[off + ai], rg[i, off + ai]
                        cx, cy = 0.5 * (an[:, 0] + an[:, 2]), 0.5 * (an[:, 1] + an[:, 3])
                        wv, hv = an[:, 2] - an[:, 0], an[:, 3] - an[:, 1]
                        bx = torch.stack((cx - r[:, 0] * wv, cy - r[:, 1] * hv, cx + r[:, 2] * wv, cy + r[:, 3] * hv), 1)
                        res.append((ops.clip_boxes_to_image(bx, (h, w)), s2, idx % 9))
                        off += nl
                return res
            us = time_op(chain, args.iters)
            out.append(row("torch_fcos_candidates (ATen chain of fcos.py:576-597)", f"B8 {n} locations", us, 8 * n * 40 + 8 * 5000 * 45))


def bench_grl_conv(args, out):
    """GRL in front of ImageDAFPN.Conv1 (DGcommon.py:73-74): stand-alone kernel + conv vs the reversal folded into the dgrad."""
    conv = torch.nn.Conv2d(256, 256, 3, stride=(2, 4)).to(DEV).to(memory_format=torch.channels_last)
    x = torch.randn(8, 256, 152, 256, device=DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    for name, fn in (("grl_conv2d fwd+bwd (fused into dgrad)", lambda: ops.grl_conv2d(x, conv)),
                     ("grad_reverse + conv fwd+bwd (separate 637 MB pass)", lambda: conv(ops.grad_reverse(x)))):
        y = fn()
        go = torch.randn_like(y)

        def step():
            x.grad = None
            fn().backward(go)
        us = time_op(step, args.iters)
        out.append(row(name, "8x256x152x256 -> Conv2d(256,256,3,stride=(2,4))", us, 2 * x.numel() * 4))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--ops", default="roi,nms,match,fcos,rpn,grl,transform,sampler,fcos_post,grl_conv")
    p.add_argument("--iters", type=int, default=20)
    p.add_argument("--batch", type=int, default=8)
    p.add_argument("--height", type=int, default=608)
    p.add_argument("--width", type=int, default=1024)
    p.add_argument("--rois", type=lambda s: [int(x) for x in s.split(",")], default=[512, 2048])
    p.add_argument("--nms", type=lambda s: [int(x) for x in s.split(",")], default=[1000, 3000, 10000, 30000, 100000])
    p.add_argument("--tv", action="store_true")
    p.add_argument("--json", default="")
    args = p.parse_args()
    out = []
    table = {"roi": bench_roi, "nms": bench_nms, "match": bench_match, "fcos": bench_fcos, "rpn": bench_rpn, "grl": bench_grl,
             "transform": bench_transform, "sampler": bench_sampler, "fcos_post": bench_fcos_post, "grl_conv": bench_grl_conv}
    for name in args.ops.split(","):
        table[name](args, out)
    if args.json:
        Path(args.json).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
