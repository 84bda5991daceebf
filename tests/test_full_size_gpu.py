"""The CUDA path at BASELINE.json's full sizes (configs[3]).  Where the C oracle still
finishes in seconds the comparison is direct and bit-exact; the rest goes through size-independent
properties: idempotence / sortedness / agreement with the stock CUDA op for NMS, linearity and the
adjoint identity <f(x), g> = <x, f^T(g)> tying MultiScaleRoIAlign's forward and backward together."""
import numpy as np
import pytest
import torch

from dgod_b200 import synth
from oracle import cpu as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from dgod_b200 import ops
    return ops


def _fpn_anchors(h, w):
    from dgod_b200.detector import grid_anchors, make_cell_anchors
    cells = make_cell_anchors(((32,), (64,), (128,), (256,), (512,)), ((0.5, 1.0, 2.0),) * 5)
    grids = [(-(-h // s), -(-w // s)) for s in (4, 8, 16, 32, 64)]
    return grid_anchors(cells, grids, [(h // gh, w // gw) for gh, gw in grids], "cpu"), grids, cells


def test_anchor_labels_at_268569_anchors_bit_exact():
    ops = _ops()
    anchors, _, _ = _fpn_anchors(800, 1344)
    assert anchors.shape[0] == 268569
    gt = [synth.random_boxes(20, 800, 1333, synth.gen(30 + i)) for i in range(2)]
    out = ops.match_boxes([g.to(DEV) for g in gt], anchors.to(DEV), 0.7, 0.3, True, want=("labels_f32", "matched_boxes"))
    for i, g in enumerate(gt):
        idx, lab, mb = O.rpn_assign(g.numpy(), anchors.numpy(), 0.7, 0.3)
        assert np.array_equal(out["matched_idx"][i].cpu().numpy(), idx)
        assert np.array_equal(out["labels_f32"][i].cpu().numpy(), lab)
        assert np.array_equal(out["matched_boxes"][i].cpu().numpy(), mb)
        assert (lab == 1).sum() >= 20          # every ground truth keeps at least its low-quality match


def test_matcher_100_gt_x_2000_proposals_bit_exact():
    ops = _ops()
    gts = [synth.random_boxes(100, 800, 1333, synth.gen(40 + i)) for i in range(8)]
    props = [synth.random_boxes(2000, 800, 1333, synth.gen(50 + i)) for i in range(8)]
    labels = [torch.randint(1, 9, (100,), generator=synth.gen(60 + i)) for i in range(8)]
    out = ops.match_boxes([g.to(DEV) for g in gts], [p.to(DEV) for p in props], 0.5, 0.5, False,
                          gt_labels=[l.to(DEV) for l in labels], want=("labels_i64", "clamped_idx"))
    off = 0
    for g, p, l in zip(gts, props, labels):
        idx, lab = O.roi_assign(g.numpy(), l.numpy(), p.numpy(), 0.5, 0.5)
        assert np.array_equal(out["labels_i64"][off:off + 2000].cpu().numpy(), lab)
        assert np.array_equal(out["clamped_idx"][off:off + 2000].cpu().numpy(), np.maximum(idx, 0))
        off += 2000


def test_fcos_assign_22400_locations_bit_exact():
    ops = _ops()
    from dgod_b200.detector import grid_anchors
    strides = (8, 16, 32, 64, 128)
    grids = [(-(-800 // s), -(-1344 // s)) for s in strides]
    cells = [torch.tensor([[-4.0 * s, -4.0 * s, 4.0 * s, 4.0 * s]]) for s in strides]
    anchors = grid_anchors(cells, grids, [(s, s) for s in strides], "cpu").float()
    npl = [gh * gw for gh, gw in grids]
    assert anchors.shape[0] == 22400
    gts = [synth.random_boxes(20, 800, 1333, synth.gen(70 + i)) for i in range(8)]
    labels = [torch.randint(1, 9, (20,), generator=synth.gen(80 + i)) for i in range(8)]
    idx = ops.fcos_assign(anchors.to(DEV), [g.to(DEV) for g in gts], npl, 1.5)
    for i, (g, l) in enumerate(zip(gts, labels)):
        ref = O.fcos_assign(anchors.numpy(), npl[0], npl[-1], g.numpy(), l.numpy())[0]
        assert np.array_equal(idx[i].cpu().numpy(), ref)


@pytest.mark.parametrize("n", [30000, 100000])
def test_batched_nms_large_properties(n):
    import torchvision
    ops = _ops()
    g = synth.gen(n)
    boxes = synth.random_boxes(n, 800, 1333, g).to(DEV)
    scores = synth.distinct_scores(n, g).to(DEV)
    idxs = torch.randint(0, 5, (n,), generator=g).to(DEV)
    keep = ops.batched_nms(boxes, scores, idxs, 0.7)
    # bit-exact against the C oracle (torchvision's CPU arithmetic: fp32 IoU compared in double) at the full size
    ref = O.batched_nms(boxes.cpu().numpy(), scores.cpu().numpy(), idxs.cpu().numpy(), 0.7, mode=0)
    assert np.array_equal(keep.cpu().numpy(), ref)
    ks = scores[keep]
    assert keep.dtype == torch.int64 and len(torch.unique(keep)) == len(keep)
    assert bool((ks[:-1] > ks[1:]).all())                                        # descending score
    again = ops.batched_nms(boxes[keep], ks, idxs[keep], 0.7)                    # idempotence: survivors do not suppress each other
    assert torch.equal(again, torch.arange(len(keep), device=DEV))
    # every suppressed box has a kept, higher-scored box of its group with IoU > 0.7 (sampled: the full
    # matrix would be 100k x 70k)
    kept_mask = torch.zeros(n, dtype=torch.bool, device=DEV)
    kept_mask[keep] = True
    supp = torch.nonzero(~kept_mask)[:, 0][:512]
    iou = ops.box_iou(boxes[supp], boxes[keep])
    ok = (iou > 0.7) & (idxs[supp][:, None] == idxs[keep][None, :]) & (ks[None, :] > scores[supp][:, None])
    assert bool(ok.any(1).all())
    # the stock CUDA op decides in fp32 (the CPU op, our oracle, in double): identical up to borderline pairs
    tv = torchvision.ops.batched_nms(boxes, scores, idxs, 0.7)
    diff = len(set(tv.tolist()) ^ set(keep.tolist()))
    assert diff <= max(4, n // 10000)


def test_msroi_align_8192_rois_linearity_and_adjoint():
    ops = _ops()
    B, C, H, W, per = 2, 256, 800, 1344, 8192
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    gx = synth.gen(5)
    xs = [torch.randn(B, C, -(-H // s), -(-W // s), generator=gx).to(DEV).contiguous(memory_format=torch.channels_last)
          for s in (4, 8, 16, 32)]
    ys = [torch.randn_like(x) for x in xs]
    boxes = [synth.random_boxes(per, H, W, synth.gen(90 + i)) for i in range(B)]
    rois = synth.rois_from_boxes(boxes).to(DEV)
    offs = ops._offsets([per] * B, DEV)
    f = lambda feats: ops.multiscale_roi_align(feats, rois, scales, 7, 2, 2, 5, roi_img_offsets=offs)
    with torch.no_grad():
        fx, fy = f(xs), f(ys)
        fz = f([2.5 * x - y for x, y in zip(xs, ys)])
    assert fx.shape == (B * per, C, 7, 7)
    torch.testing.assert_close(fz, 2.5 * fx - fy, rtol=1e-4, atol=1e-4)
    # adjoint identity: <f(x), g> == sum_l <x_l, grad_l>  (the backward is the transpose of the forward)
    xg = [x.clone().requires_grad_(True) for x in xs]
    out = f(xg)
    go = torch.randn(out.shape, generator=torch.Generator(device=DEV).manual_seed(3), device=DEV)
    out.backward(go)
    lhs = float((out.detach().double() * go.double()).sum())
    rhs = sum(float((x.detach().double() * x.grad.double()).sum()) for x in xg)
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), float((out.detach().double().abs() * go.double().abs()).sum()) * 1e-2)
    # a sample of RoIs against the oracle (the op itself at this size takes the oracle minutes)
    sel = torch.arange(0, B * per, 997)
    ref = O.msroi_align_fwd([x.detach().cpu().contiguous().numpy() for x in xs], rois[sel].cpu().numpy(), scales, 7, 7, 2, 2, 5)
    np.testing.assert_allclose(fx[sel].cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())


def test_msroi_align_repeated_launches_agree():
    """The persistent backward (driver lane / row_full mbarriers / rows_released, zero-fill warp, bulk reductions) and
    the multi-lane forward at bench.py's shape, 25 launches back to back with a dirty output buffer in between: every
    result equals the first up to the reduction order (an intermittent hand-off race would show as a large error)."""
    ops = _ops()
    B, C, H, W, per = 8, 256, 608, 1024, 512
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    xs = [f.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in synth.random_features(B, C, H, W, 0)]
    rois = synth.rois_from_boxes([synth.random_boxes(per, H, W, synth.gen(10 + i)) for i in range(B)]).to(DEV)
    offs = ops._offsets([per] * B, DEV)
    go = torch.randn(B * per, C, 7, 7, generator=synth.gen(99)).to(DEV)
    first_out = first_grads = None
    for it in range(25):
        out = ops.multiscale_roi_align(xs, rois, scales, 7, 2, 2, 5, roi_img_offsets=offs)
        for x in xs:
            x.grad = None
        out.backward(go)
        if first_out is None:
            first_out, first_grads = out.detach().clone(), [x.grad.clone() for x in xs]
            scale = [float(g.abs().max()) for g in first_grads]
            continue
        assert torch.equal(out.detach(), first_out)                      # the forward is deterministic
        for x, g0, s in zip(xs, first_grads, scale):
            assert float((x.grad - g0).abs().max()) <= 2e-6 * s
