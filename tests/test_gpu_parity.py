"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.
Bit-exact for integer / index work; RoIAlign within 1e-5 (fp32) / 1e-2 (bf16)."""
import numpy as np
import pytest
import torch

from dgod_b200 import synth
from oracle import cpu as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _boxes(n, seed, h=800, w=1333):
    return synth.random_boxes(n, h, w, synth.gen(seed))


def _ops():
    from dgod_b200 import ops
    return ops


# ------------------------------------------------------------------------------- box_iou / Matcher
@pytest.mark.parametrize("m,n", [(1, 1), (20, 777), (100, 2000)])
def test_box_iou_bit_exact(m, n):
    a, b = _boxes(m, 0), _boxes(n, 1)
    got = _ops().box_iou(a.to(DEV), b.to(DEV)).cpu().numpy()
    ref = O.box_iou(a.numpy(), b.numpy())
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("m,n,hi,lo,lq", [(20, 5000, 0.7, 0.3, True), (100, 2000, 0.5, 0.5, False), (1, 300, 0.7, 0.3, True), (300, 999, 0.6, 0.2, True)])
def test_matcher_on_matrix(m, n, hi, lo, lq):
    ops = _ops()
    q = O.box_iou(_boxes(m, 2).numpy(), _boxes(n, 3).numpy())
    got = ops.Matcher(hi, lo, lq)(torch.from_numpy(q).to(DEV)).cpu().numpy()
    assert np.array_equal(got, O.matcher(q, hi, lo, lq))


def test_matcher_errors_and_edges():
    ops = _ops()
    with pytest.raises(ValueError, match="No ground-truth"):
        ops.Matcher(0.5, 0.5)(torch.zeros((0, 5), device=DEV))
    with pytest.raises(ValueError, match="No proposal"):
        ops.Matcher(0.5, 0.5)(torch.zeros((5, 0), device=DEV))
    q = np.array([[np.float32(0.7), np.float32(0.3), 0.29999998, 0.0, 0.2, 0.2],
                  [0.0, 0.0, 0.0, 0.0, 0.0, 0.0]], dtype=np.float32)
    got = ops.Matcher(0.7, 0.3, True)(torch.from_numpy(q).to(DEV)).cpu().numpy()
    assert np.array_equal(got, O.matcher(q, 0.7, 0.3, True))


@pytest.mark.parametrize("gts", [[20, 20], [20, 0, 1, 300], [5]])
def test_fused_rpn_labels(gts):
    ops = _ops()
    anchors = _boxes(50000, 4, 608, 1024)
    gt = [_boxes(m, 10 + i, 608, 1024) for i, m in enumerate(gts)]
    out = ops.match_boxes([g.to(DEV) for g in gt], anchors.to(DEV), 0.7, 0.3, True,
                          want=("labels_f32", "matched_boxes"))
    for i, g in enumerate(gt):
        idx, lab, mb = O.rpn_assign(g.numpy(), anchors.numpy(), 0.7, 0.3)
        assert np.array_equal(out["matched_idx"][i].cpu().numpy(), idx)
        assert np.array_equal(out["labels_f32"][i].cpu().numpy(), lab)
        assert np.array_equal(out["matched_boxes"][i].cpu().numpy(), mb)


def test_fused_roi_labels():
    ops = _ops()
    gts, props = [20, 0, 3], [2020, 1500, 2003]
    gt = [_boxes(m, 20 + i, 608, 1024) for i, m in enumerate(gts)]
    lab = [torch.randint(1, 9, (m,), generator=synth.gen(30 + i)) for i, m in enumerate(gts)]
    pr = [torch.cat([_boxes(n - m, 40 + i, 608, 1024), g]) for i, (n, m, g) in enumerate(zip(props, gts, gt))]
    out = ops.match_boxes([g.to(DEV) for g in gt], [p.to(DEV) for p in pr], 0.5, 0.5, False,
                          gt_labels=[l.to(DEV) for l in lab], want=("labels_i64", "clamped_idx"))
    off = 0
    for g, l, p in zip(gt, lab, pr):
        ci, lb = O.roi_assign(g.numpy(), l.numpy(), p.numpy(), 0.5, 0.5)
        assert np.array_equal(out["clamped_idx"][off:off + len(p)].cpu().numpy(), ci)
        assert np.array_equal(out["labels_i64"][off:off + len(p)].cpu().numpy(), lb)
        off += len(p)


# ------------------------------------------------------------------------------- FCOS
def _fcos_anchors(h, w):
    out, npl = [], []
    for s in (8, 16, 32, 64, 128):
        gh, gw = -(-h // s), -(-w // s)
        cell = np.array([[-s * 4, -s * 4, s * 4, s * 4]], np.float32)  # anchor size 8*stride (fcos.py:467-469)
        out.append(O.grid_anchors(cell, gh, gw, s, s))
        npl.append(gh * gw)
    return np.concatenate(out), npl


@pytest.mark.parametrize("gts", [[20, 20], [0, 1, 2, 20], [300]])
def test_fcos_assign_bit_exact(gts):
    ops = _ops()
    anchors, npl = _fcos_anchors(608, 1024)
    gt = [_boxes(m, 50 + i, 608, 1024) for i, m in enumerate(gts)]
    # equal quirky areas (fcos.py:543) force real ties in 1e8 - area
    if gts[0] >= 2:
        gt[0][1] = gt[0][0] + torch.tensor([3.0, 0.0, 3.0, 0.0])
    lab = [torch.randint(1, 9, (m,), generator=synth.gen(60 + i)) for i, m in enumerate(gts)]
    idx, cls, bt, oh = ops.fcos_assign(torch.from_numpy(anchors).to(DEV), [g.to(DEV) for g in gt], npl, 1.5,
                                       gt_labels=[l.to(DEV) for l in lab], num_classes=9)
    plain = ops.fcos_assign(torch.from_numpy(anchors).to(DEV), [g.to(DEV) for g in gt], npl, 1.5)
    assert torch.equal(plain, idx)
    for i, (g, l) in enumerate(zip(gt, lab)):
        ri, rc, rb = O.fcos_assign(anchors, npl[0], npl[-1], g.numpy(), l.numpy(), 1.5)
        assert np.array_equal(idx[i].cpu().numpy(), ri)
        assert np.array_equal(cls[i].cpu().numpy(), rc)
        assert np.array_equal(bt[i].cpu().numpy(), rb)
        onehot = np.zeros((len(ri), 9), np.float32)
        fg = rc >= 0
        onehot[np.nonzero(fg)[0], rc[fg]] = 1.0
        assert np.array_equal(oh[i].cpu().numpy(), onehot)
    assert (idx.cpu().numpy() >= 0).sum() > 0


# ------------------------------------------------------------------------------- NMS
@pytest.mark.parametrize("n,thr", [(1, 0.5), (2, 0.5), (63, 0.5), (64, 0.3), (65, 0.7), (1000, 0.7), (4097, 0.6), (9000, 0.7), (20000, 0.5)])
def test_nms_bit_exact(n, thr):
    boxes = _boxes(n, n)
    scores = torch.rand(n, generator=synth.gen(n + 1))  # fp32 rand: ties appear for large n (stable order)
    got = _ops().nms(boxes.to(DEV), scores.to(DEV), thr).cpu().numpy()
    assert np.array_equal(got, O.nms(boxes.numpy(), scores.numpy(), thr))


def test_nms_edges():
    ops = _ops()
    assert ops.nms(torch.zeros((0, 4), device=DEV), torch.zeros(0, device=DEV), 0.5).shape == (0,)
    b = torch.tensor([[0, 0, 10, 10], [20, 20, 30, 30], [40, 40, 50, 50], [60, 60, 70, 70]], dtype=torch.float32)
    for s in ([0.5, 0.5, 0.5, 0.5], [0.3, 0.9, 0.3, 0.9]):
        s = torch.tensor(s)
        assert ops.nms(b.to(DEV), s.to(DEV), 0.5).tolist() == O.nms(b.numpy(), s.numpy(), 0.5).tolist()
    s = torch.tensor([0.9, 0.8])
    for bb, thr in (([[0, 0, 2, 1], [0, 0, 1, 1]], 0.5), ([[0, 0, 5, 1], [0, 0, 3, 1]], 0.6), ([[0, 0, 10, 1], [0, 0, 7, 1]], 0.7)):
        bb = torch.tensor(bb, dtype=torch.float32)
        assert ops.nms(bb.to(DEV), s.to(DEV), thr).tolist() == O.nms(bb.numpy(), s.numpy(), thr).tolist()
    dup = torch.tensor([[0, 0, 5, 5]] * 200, dtype=torch.float32)
    sc = torch.linspace(0, 1, 200)
    assert ops.nms(dup.to(DEV), sc.to(DEV), 0.5).tolist() == [199]
    assert ops.nms(dup.to(DEV), torch.ones(200, device=DEV), 0.5).tolist() == [0]


@pytest.mark.parametrize("n,groups,thr", [(300, 3, 0.5), (1000, 5, 0.7), (1001, 5, 0.7), (4096, 8, 0.5), (5000, 9, 0.6), (8819, 5, 0.7), (30000, 5, 0.7)])
def test_batched_nms_bit_exact(n, groups, thr):
    g = synth.gen(n)
    boxes = _boxes(n, n + 7, 608, 1024)
    scores = synth.distinct_scores(n, g)
    idxs = torch.randint(0, groups, (n,), generator=g)
    got = _ops().batched_nms(boxes.to(DEV), scores.to(DEV), idxs.to(DEV), thr).cpu().numpy()
    assert np.array_equal(got, O.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), thr))


def test_batched_nms_ties_large_group_ids_and_modes():
    ops = _ops()
    g = synth.gen(3)
    n = 3000
    boxes = _boxes(n, 5)
    scores = torch.randint(0, 40, (n,), generator=g).float() / 40
    idxs = torch.randint(0, 5, (n,), generator=g) * 1_000_003 - 17      # sparse / negative ids
    got = ops.batched_nms(boxes.to(DEV), scores.to(DEV), idxs.to(DEV), 0.7).cpu().numpy()
    assert np.array_equal(got, O.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), 0.7))
    # coordinate-trick arithmetic (<= 1000 boxes) incl. its fp32 rounding
    n = 900
    boxes = _boxes(n, 6) * 37.3
    scores = synth.distinct_scores(n, g)
    idxs = torch.randint(0, 90, (n,), generator=g)
    got = ops.batched_nms(boxes.to(DEV), scores.to(DEV), idxs.to(DEV), 0.5).cpu().numpy()
    assert np.array_equal(got, O.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), 0.5, mode=1))
    assert ops.batched_nms(torch.zeros((0, 4), device=DEV), torch.zeros(0, device=DEV),
                           torch.zeros(0, dtype=torch.int64, device=DEV), 0.5).shape == (0,)


def test_nms_segments_batch_with_valid_mask_and_truncation():
    ops = _ops()
    counts = [4096, 0, 1, 3333, 700]
    g = synth.gen(11)
    boxes = torch.cat([_boxes(c, 70 + i, 608, 1024) for i, c in enumerate(counts)])
    n = boxes.shape[0]
    scores = synth.distinct_scores(n, g)
    groups = torch.randint(1, 9, (n,), generator=g)
    valid = torch.rand(n, generator=g) > 0.3
    keep, info = ops.nms_segments(boxes.to(DEV), scores.to(DEV), groups.to(DEV), counts, 0.5,
                                  valid=valid.to(DEV), max_out_per_seg=100)
    keep, info = keep.cpu().numpy(), info.cpu().numpy()
    assert info[-1] == 0
    off = 0
    for s, c in enumerate(counts):
        sl = slice(off, off + c)
        vi = np.nonzero(valid[sl].numpy())[0]
        ref = O.batched_nms(boxes[sl].numpy()[vi], scores[sl].numpy()[vi], groups[sl].numpy()[vi], 0.5, mode=0)
        ref = vi[ref][:100]
        assert info[s] == len(ref)
        assert np.array_equal(keep[s, :len(ref)], ref)
        off += c


# ------------------------------------------------------------------------------- RoIAlign
def _roi_cases(H, W, scale):
    img_h, img_w = H / scale, W / scale
    return torch.tensor([
        [0, 10.3, 20.7, 200.2, 180.9], [1, 0, 0, img_w, img_h], [0, 50, 50, 50.2, 50.3],
        [1, -30, -40, 60, 70], [0, img_w - 20, img_h - 20, img_w + 90, img_h + 80],
        [1, img_w + 40, img_h + 40, img_w + 80, img_h + 90], [0, 100, 100, 100, 100]], dtype=torch.float32)


@pytest.mark.parametrize("sr,aligned,scale,out", [(2, False, 0.25, 7), (2, False, 0.125, 7), (0, False, 0.25, 7), (2, True, 0.25, 7), (3, False, 0.0625, (5, 9)), (2, False, 0.25, 14)])
@pytest.mark.parametrize("nhwc", [False, True])
def test_roi_align_forward_backward(sr, aligned, scale, out, nhwc):
    ops = _ops()
    H, W, C = 48, 64, 40
    x = torch.randn(2, C, H, W, generator=synth.gen(0))
    rois = torch.cat([_roi_cases(H, W, scale), synth.rois_from_boxes([_boxes(40, 1, H / scale, W / scale), _boxes(40, 2, H / scale, W / scale)])])
    ph, pw = (out, out) if isinstance(out, int) else out
    ref = O.roi_align_fwd(x.numpy(), rois.numpy(), scale, ph, pw, sr, aligned)
    xd = x.to(DEV)
    if nhwc:
        xd = xd.contiguous(memory_format=torch.channels_last)
    xd.requires_grad_(True)
    got = ops.roi_align(xd, rois.to(DEV), out, scale, sr, aligned)
    np.testing.assert_allclose(got.detach().cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    go = torch.randn(ref.shape, generator=synth.gen(5))
    got.backward(go.to(DEV))
    gref = O.roi_align_bwd(go.numpy(), tuple(x.shape), rois.numpy(), scale, sr, aligned)
    np.testing.assert_allclose(xd.grad.cpu().numpy(), gref, rtol=1e-5, atol=1e-5 * np.abs(gref).max())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("nhwc", [False, True])
def test_multiscale_roi_align(dtype, tol, nhwc):
    from dgod_b200.poolers import MultiScaleRoIAlign
    img_h, img_w, C = 320, 416, 64
    feats = synth.random_features(2, C, img_h, img_w, seed=3)
    if dtype == torch.bfloat16:
        feats = [f.to(dtype).float() for f in feats]     # oracle = fp32 on bf16-rounded inputs
    sides = [111.99, 112.0, 112.01, 223.99, 224.0, 224.01, 447.9, 448.0, 448.1]
    edge = torch.tensor([[3.0, 5.0, 3.0 + s, 5.0 + s] for s in sides])
    boxes = [torch.cat([_boxes(150, 5, img_h, img_w), edge]), _boxes(130, 6, img_h, img_w)]
    rois = synth.rois_from_boxes(boxes)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    ref = O.msroi_align_fwd([f.numpy() for f in feats], rois.numpy(), scales, 7, 7, 2, 2, 5)
    pool = MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2)
    x = {}
    for i, f in enumerate(feats):
        t = f.to(DEV).to(dtype)
        if nhwc:
            t = t.contiguous(memory_format=torch.channels_last)
        x[str(i)] = t.requires_grad_(True)
    x["pool"] = torch.zeros(2, C, 5, 7, device=DEV, dtype=dtype)
    got = pool(x, [b.to(DEV) for b in boxes], [(img_h, img_w)] * 2)
    assert got.dtype == dtype and got.shape == ref.shape
    scale_ref = np.abs(ref).max()
    np.testing.assert_allclose(got.detach().float().cpu().numpy(), ref, rtol=tol, atol=tol * scale_ref)
    assert pool.scales == scales
    go = torch.randn(ref.shape, generator=synth.gen(9))
    if dtype == torch.bfloat16:
        go = go.to(dtype).float()
    grads = O.msroi_align_bwd(go.numpy(), [tuple(f.shape) for f in feats], rois.numpy(), scales, 2, 2, 5)
    from dgod_b200 import ops
    # atomic scatter / tile gather / TMA bulk reduce, owner-computes with dealt and with claimed work items (channels_last only) / auto
    algos = ((1, 2) if dtype == torch.float32 else (2,)) + ((3, 4, 5) if nhwc else ()) + (0,)
    for algo in algos:
        ops.BACKWARD_ALGO = algo
        try:
            for t in x.values():
                t.grad = None
            got.backward(go.to(DEV).to(dtype), retain_graph=True)
        finally:
            ops.BACKWARD_ALGO = 0
        for i, gr in enumerate(grads):
            gg = x[str(i)].grad
            assert gg.dtype == dtype and gg.shape == gr.shape
            assert gg.is_contiguous(memory_format=torch.channels_last if nhwc else torch.contiguous_format)
            np.testing.assert_allclose(gg.float().cpu().numpy(), gr, rtol=tol, atol=tol * max(np.abs(gr).max(), 1e-3))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("with_offsets", [True, False])
def test_multiscale_roi_align_tma_paths(dtype, tol, with_offsets):
    """The TMA kernels (channels_last, C = 256, 7x7, sr 2) on the shapes that stress them: footprints
    wider / taller than one 32-pixel chunk (whole-image RoIs on every level), sub-pixel RoIs, RoIs
    hanging over every border, an RoI completely outside, and the per-image persistent schedule
    (roi_img_offsets) vs the plain one."""
    from dgod_b200 import ops
    img_h, img_w, C, B = 352, 1344, 256, 3
    feats = synth.random_features(B, C, img_h, img_w, seed=11)
    if dtype == torch.bfloat16:
        feats = [f.to(dtype).float() for f in feats]
    special = torch.tensor([[0.0, 0.0, img_w, img_h], [-50.0, -60.0, img_w + 70, img_h + 80], [7.0, 9.0, 7.4, 9.3],
                            [img_w - 3.0, 2.0, img_w + 40, 30.0], [img_w + 10.0, img_h + 10.0, img_w + 60, img_h + 90],
                            [5.0, 5.0, 1300.0, 40.0], [20.0, 3.0, 60.0, 349.0], [100.0, 100.0, 100.0, 100.0]])
    boxes = [torch.cat([_boxes(200, 20 + i, img_h, img_w), special]) for i in range(B)]
    rois = synth.rois_from_boxes(boxes)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    ref = O.msroi_align_fwd([f.numpy() for f in feats], rois.numpy(), scales, 7, 7, 2, 2, 5)
    xs = [f.to(DEV).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in feats]
    offs = ops._offsets([b.shape[0] for b in boxes], DEV) if with_offsets else None
    got = ops.multiscale_roi_align(xs, rois.to(DEV), scales, 7, 2, 2, 5, roi_img_offsets=offs)
    np.testing.assert_allclose(got.detach().float().cpu().numpy(), ref, rtol=tol, atol=tol * np.abs(ref).max())
    go = torch.randn(ref.shape, generator=synth.gen(12))
    if dtype == torch.bfloat16:
        go = go.to(dtype).float()
    grads = O.msroi_align_bwd(go.numpy(), [tuple(f.shape) for f in feats], rois.numpy(), scales, 2, 2, 5)
    ops.BACKWARD_ALGO = 3
    try:
        for rep in range(2):        # twice: the outputs must be overwritten, not accumulated
            for t in xs:
                t.grad = None
            got.backward(go.to(DEV).to(dtype), retain_graph=True)
    finally:
        ops.BACKWARD_ALGO = 0
    for x, gr in zip(xs, grads):
        np.testing.assert_allclose(x.grad.float().cpu().numpy(), gr, rtol=tol, atol=tol * max(np.abs(gr).max(), 1e-3))
    # size-independent property: every in-range sample spreads weight 1/count, so the gradient of
    # sum(out) sums to (#valid samples / count) per channel == sum of the oracle's gradient
    tot = sum(float(x.grad.float().sum()) for x in xs)
    np.testing.assert_allclose(tot, sum(float(g.astype(np.float64).sum()) for g in grads), rtol=5e-3 if dtype == torch.bfloat16 else 1e-4)


@pytest.mark.parametrize("C,sr", [(64, 2), (128, 2), (256, 1)])
def test_multiscale_roi_align_tma_other_instantiations(C, sr):
    """The other instantiated shapes of the TMA kernels (1 or 2 consumer warps, sampling_ratio 1), a single
    RoI, and zero RoIs (the backward must still overwrite the gradient maps with zeros)."""
    from dgod_b200 import ops
    img_h, img_w, B = 200, 264, 2
    feats = synth.random_features(B, C, img_h, img_w, seed=31)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    for n in (150, 1, 0):
        boxes = [_boxes(n, 40 + i, img_h, img_w) for i in range(B)]
        rois = synth.rois_from_boxes(boxes).reshape(-1, 5)
        xs = [f.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in feats]
        got = ops.multiscale_roi_align(xs, rois.to(DEV), scales, 7, sr, 2, 5)
        assert got.shape == (rois.shape[0], C, 7, 7)
        go = torch.randn(got.shape, generator=synth.gen(32))
        ops.BACKWARD_ALGO = 3
        try:
            for t in xs:
                t.grad = torch.full_like(t, 7.0).contiguous(memory_format=torch.channels_last)   # must not leak through
                t.grad = None
            got.backward(go.to(DEV))
        finally:
            ops.BACKWARD_ALGO = 0
        if n == 0:
            assert all(float(x.grad.abs().max()) == 0.0 for x in xs)
            continue
        ref = O.msroi_align_fwd([f.numpy() for f in feats], rois.numpy(), scales, 7, 7, sr, 2, 5)
        np.testing.assert_allclose(got.detach().cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
        grads = O.msroi_align_bwd(go.numpy(), [tuple(f.shape) for f in feats], rois.numpy(), scales, sr, 2, 5)
        for x, gr in zip(xs, grads):
            np.testing.assert_allclose(x.grad.cpu().numpy(), gr, rtol=1e-5, atol=1e-5 * max(np.abs(gr).max(), 1e-3))


def test_roi_align_empty_and_errors():
    ops = _ops()
    x = torch.randn(1, 8, 16, 16, device=DEV)
    assert ops.roi_align(x, torch.zeros((0, 5), device=DEV), 7, 0.25, 2).shape == (0, 8, 7, 7)
    with pytest.raises(RuntimeError, match="Tensor\\[K, 5\\]"):
        ops.roi_align(x, torch.zeros((3, 4), device=DEV), 7)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.roi_align(x.cpu(), torch.zeros((1, 5)), 7)


# ------------------------------------------------------------------------------- RPN pipeline
def _rpn_case(B, img_h, img_w, seed, pre=600):
    from torchvision.models.detection.anchor_utils import AnchorGenerator
    g = synth.gen(seed)
    ag = AnchorGenerator(((32,), (64,), (128,), (256,), (512,)), ((0.5, 1.0, 2.0),) * 5)
    ph, pw = -(-img_h // 32) * 32, -(-img_w // 32) * 32
    grids = [(-(-ph // s), -(-pw // s)) for s in (4, 8, 16, 32, 64)]
    strides = [(ph // h, pw // w) for h, w in grids]
    obj = [torch.randn(B, 3, h, w, generator=g) for h, w in grids]
    dl = [torch.randn(B, 12, h, w, generator=g) * 0.3 for h, w in grids]
    cell = [c.tolist() for c in ag.cell_anchors]
    return obj, dl, grids, strides, cell


def _flatten_like_torchvision(obj, dl):
    B = obj[0].shape[0]
    o = torch.cat([x.permute(0, 2, 3, 1).reshape(B, -1) for x in obj], 1)
    d = torch.cat([x.view(B, -1, 4, x.shape[2], x.shape[3]).permute(0, 3, 4, 1, 2).reshape(B, -1, 4) for x in dl], 1)
    return o, d


@pytest.mark.parametrize("B,img,pre,post", [(2, (300, 500), 600, 500), (3, (600, 999), 2000, 2000)])
def test_rpn_proposals_fused(B, img, pre, post):
    ops = _ops()
    obj, dl, grids, strides, cell = _rpn_case(B, img[0], img[1], 5, pre)
    sizes = torch.tensor([[img[0], img[1]]] * B, dtype=torch.float32)
    boxes, scores, counts = ops.rpn_proposals([o.to(DEV) for o in obj], [d.to(DEV) for d in dl], sizes.to(DEV),
                                              strides, cell, pre, post, 0.7, 1e-3, 0.0)
    boxes, scores, counts = boxes.cpu().numpy(), scores.cpu().numpy(), counts.cpu().numpy()
    o_flat, d_flat = _flatten_like_torchvision(obj, dl)
    anchors = np.concatenate([O.grid_anchors(np.array(cell[l], np.float32), h, w, strides[l][0], strides[l][1])
                              for l, (h, w) in enumerate(grids)])
    npl = [3 * h * w for h, w in grids]
    for i in range(B):
        props = O.box_decode(d_flat[i].numpy(), anchors)
        rb, rs = O.rpn_filter_image(props, o_flat[i].numpy(), npl, pre, post, 0.7, 1e-3, 0.0, img[0], img[1])
        assert counts[i] == len(rb)
        assert np.array_equal(boxes[i, :len(rb)], rb)       # same exp/sigmoid definition -> bit-exact
        assert np.array_equal(scores[i, :len(rb)], rs)
        assert not boxes[i, len(rb):].any()


def test_rpn_filter_on_decoded_proposals():
    ops = _ops()
    g = synth.gen(4)
    npl = [3 * 38 * 64, 3 * 19 * 32, 3 * 10 * 16, 3 * 5 * 8, 3 * 3 * 4]
    A = sum(npl)
    props = torch.stack([synth.random_boxes(A, 330, 540, g), synth.random_boxes(A, 330, 540, g)]) - 15.0
    obj = torch.randn(2, A, generator=g)
    sizes = torch.tensor([[300.0, 500.0], [280.0, 500.0]])
    boxes, scores, counts = ops.rpn_filter_proposals(props.to(DEV), obj.to(DEV), sizes.to(DEV), npl, 600, 500, 0.7)
    for i in range(2):
        rb, rs = O.rpn_filter_image(props[i].numpy(), obj[i].numpy(), npl, 600, 500, 0.7, 1e-3, 0.0,
                                    float(sizes[i, 0]), float(sizes[i, 1]))
        assert int(counts[i]) == len(rb)
        assert np.array_equal(boxes[i, :len(rb)].cpu().numpy(), rb)
        assert np.array_equal(scores[i, :len(rb)].cpu().numpy(), rs)


def test_rpn_topk_with_massive_ties():
    ops = _ops()
    obj, dl, grids, strides, cell = _rpn_case(1, 300, 500, 9)
    obj = [torch.zeros_like(o) for o in obj]            # every logit equal: ties resolved by anchor index
    obj[0][0, 1, 3, 5] = 1.0
    sizes = torch.tensor([[300.0, 500.0]])
    boxes, scores, counts = ops.rpn_proposals([o.to(DEV) for o in obj], [d.to(DEV) for d in dl], sizes.to(DEV),
                                              strides, cell, 300, 200, 0.7)
    o_flat, d_flat = _flatten_like_torchvision(obj, dl)
    anchors = np.concatenate([O.grid_anchors(np.array(cell[l], np.float32), h, w, strides[l][0], strides[l][1])
                              for l, (h, w) in enumerate(grids)])
    rb, rs = O.rpn_filter_image(O.box_decode(d_flat[0].numpy(), anchors), o_flat[0].numpy(),
                                [3 * h * w for h, w in grids], 300, 200, 0.7, 1e-3, 0.0, 300, 500)
    assert int(counts[0]) == len(rb)
    assert np.array_equal(boxes[0, :len(rb)].cpu().numpy(), rb)


# ------------------------------------------------------------------------------- post-processing, GRL
def test_detect_candidates_decode_and_grl():
    ops = _ops()
    g = synth.gen(8)
    per = [512, 300]
    n, ncls = sum(per), 9
    props = synth.random_boxes(n, 600, 1000, g)
    logits, reg = torch.randn(n, ncls, generator=g) * 2, torch.randn(n, ncls * 4, generator=g) * 0.5
    sizes = torch.tensor([[600.0, 1000.0], [580.0, 990.0]])
    cb, cs, cl, cv = ops.detect_candidates(logits.to(DEV), reg.to(DEV), props.to(DEV), per, sizes.to(DEV))
    off = 0
    for i, c in enumerate(per):
        rb, rs, rl, rv = O.detect_candidates(logits[off:off + c].numpy(), reg[off:off + c].numpy(),
                                             props[off:off + c].numpy(), float(sizes[i, 0]), float(sizes[i, 1]))
        assert np.array_equal(cb[off:off + c].cpu().numpy(), rb)
        assert np.array_equal(cs[off:off + c].cpu().numpy(), rs)
        assert np.array_equal(cl[off:off + c].cpu().numpy(), rl)
        assert np.array_equal(cv[off:off + c].cpu().numpy(), rv)
        off += c
    dec = ops.box_decode(reg.to(DEV), props.to(DEV), (10.0, 10.0, 5.0, 5.0)).cpu().numpy()
    assert np.array_equal(dec, O.box_decode(reg.numpy(), props.numpy(), (10.0, 10.0, 5.0, 5.0)))
    for shape in [(1000,), (513, 1024), (7,)]:
        x = torch.randn(*shape, generator=g)
        xd = x.to(DEV).requires_grad_(True)
        y = ops.grad_reverse(xd)
        assert torch.equal(y, xd)
        go = torch.randn(*shape, generator=g)
        y.backward(go.to(DEV))
        assert np.array_equal(xd.grad.cpu().numpy(), O.grl_scale(go.numpy(), 0.1))
    xb = torch.randn(4097, generator=g).to(torch.bfloat16)
    xd = xb.to(DEV).requires_grad_(True)
    ops.grad_reverse(xd).backward(xb.to(DEV))
    assert torch.equal(xd.grad.cpu(), (xb.neg() * 0.1))


def test_grl_linear_matches_unfused():
    ops = _ops()
    g = synth.gen(12)
    x = torch.randn(1024, 1024, generator=g).to(DEV).requires_grad_(True)
    lin = torch.nn.Linear(1024, 512).to(DEV)
    go = torch.randn(1024, 512, generator=g).to(DEV)
    y1 = ops.grl_linear(x, lin.weight, lin.bias)
    y1.backward(go)
    g1, gw1 = x.grad.clone(), lin.weight.grad.clone()
    x.grad = None
    lin.weight.grad = None
    y2 = lin(ops.grad_reverse(x))
    y2.backward(go)
    assert torch.equal(y1, y2)
    torch.testing.assert_close(g1, x.grad, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(gw1, lin.weight.grad, rtol=1e-5, atol=1e-5)


def test_fcos_eval_candidates_kernel():
    """ops.fcos_candidates (fcos.py:576-597 in one launch) against the C oracle and, with the segmented NMS behind it,
    against the detections of the reference's own FCOS.postprocess_detections (tests/golden/fcos_post.npz)."""
    from pathlib import Path
    from oracle import gen_golden as G
    ops = _ops()
    gold = np.load(Path(__file__).parent / "golden" / "fcos_post.npz")
    anchors, npl, _, _ = G.fcos_inputs()
    ho, shapes = G.fcos_post_inputs()
    sizes = torch.tensor([[float(h), float(w)] for h, w in shapes], device=DEV)
    topk = 300
    boxes, scores, labels, valid, counts = ops.fcos_candidates(ho["cls_logits"].to(DEV), ho["bbox_regression"].to(DEV),
                                                                ho["bbox_ctrness"].to(DEV), torch.from_numpy(anchors).to(DEV),
                                                                npl, sizes, 0.2, topk)
    L = len(npl)
    for i, (h, w) in enumerate(shapes):
        rb, rs, rl, rc = O.fcos_candidates(ho["cls_logits"][i].numpy(), ho["bbox_regression"][i].numpy(),
                                           ho["bbox_ctrness"][i].numpy(), anchors, npl, h, w, 0.2, topk)
        assert np.array_equal(counts[i].cpu().numpy(), rc)
        assert np.array_equal(valid[i].cpu().numpy().reshape(L, topk), np.arange(topk)[None, :] < rc[:, None])
        got_s, got_l, got_b = scores[i].cpu().numpy(), labels[i].cpu().numpy(), boxes[i].cpu().numpy()
        np.testing.assert_allclose(got_s, rs, rtol=1e-6)                         # expf: scores to ~1 ulp
        same = got_l == rl                                                        # order can differ only where scores are 1 ulp apart
        assert same.mean() > 0.995
        assert np.array_equal(got_b[same], rb[same])
    per = L * topk
    keep, info = ops.nms_segments(boxes.view(-1, 4), scores.view(-1), labels.view(-1), [per] * len(shapes), 0.6,
                                  valid=valid.view(-1), max_out_per_seg=100)
    nums = info[:-1].tolist()
    for i in range(len(shapes)):
        k = keep[i, :nums[i]]
        assert nums[i] == len(gold[f"labels{i}"])
        assert np.array_equal(labels[i][k].cpu().numpy(), gold[f"labels{i}"])
        np.testing.assert_allclose(scores[i][k].cpu().numpy(), gold[f"scores{i}"], rtol=1e-6)
        np.testing.assert_allclose(boxes[i][k].cpu().numpy(), gold[f"boxes{i}"], rtol=0, atol=1e-4)
    # degenerate inputs: nothing passes the threshold; threshold 0 with fewer elements than topk
    none = ops.fcos_candidates(torch.full((1, len(anchors), 9), -20.0, device=DEV), ho["bbox_regression"][:1].to(DEV),
                               ho["bbox_ctrness"][:1].to(DEV), torch.from_numpy(anchors).to(DEV), npl, sizes[:1], 0.2, topk)
    assert int(none[4].sum()) == 0 and int(none[3].sum()) == 0 and float(none[0].abs().sum()) == 0.0
    few = ops.fcos_candidates(ho["cls_logits"][:1].to(DEV), ho["bbox_regression"][:1].to(DEV), ho["bbox_ctrness"][:1].to(DEV),
                              torch.from_numpy(anchors).to(DEV), npl, sizes[:1], 0.0, 1000)
    assert few[4][0].cpu().tolist() == [min(1000, 9 * n) for n in npl]


@pytest.mark.parametrize("n,dtype", [(155520, torch.float32), (2020, torch.int64), (300, torch.int64), (40, torch.float32)])
def test_balanced_sampler_kernel(n, dtype):
    """ops.balanced_sample (TV _utils.py:11-71 with explicit keys) against the numpy oracle: bit-exact index lists, incl.
    images with fewer positives / negatives than asked for, no positives at all, massive key ties, and N < batch."""
    ops = _ops()
    g = synth.gen(n)
    B = 5
    lab = torch.full((B, n), -1.0)
    r = torch.rand(B, n, generator=g)
    lab[r < 0.6] = 0.0
    lab[r < 0.01] = 1.0                       # ~1 % positives
    lab[1][lab[1] == 1] = 0.0                 # image 1: no positives
    lab[2][:] = -1.0
    lab[2][: min(n, 37)] = 1.0                # image 2: 37 positives, no negatives
    lab[3][lab[3] == 0] = -1.0                # image 3: no negatives
    keys = torch.rand(B, n, generator=g)
    keys[4] = (keys[4] * 4).floor() / 4       # image 4: only 4 distinct keys -> ties decide
    labels = lab.to(dtype) * (3 if dtype == torch.int64 else 1)
    P, S = (128, 256) if dtype == torch.float32 else (128, 512)
    pi, pv, ni, nv, cnt = ops.balanced_sample(labels.to(DEV), keys.to(DEV), P, S)
    ref = O.balanced_sample(lab.numpy(), keys.numpy(), P, S)
    for b, (rp, rn) in enumerate(ref):
        assert cnt[b].tolist() == [len(rp), len(rn)]
        assert np.array_equal(pi[b, :len(rp)].cpu().numpy(), rp) and np.array_equal(ni[b, :len(rn)].cpu().numpy(), rn)
        assert pv[b].cpu().numpy().tolist() == [i < len(rp) for i in range(pi.shape[1])]
        assert nv[b].cpu().numpy().tolist() == [i < len(rn) for i in range(ni.shape[1])]
    assert pi.shape == (B, min(P, n)) and ni.shape == (B, min(S, n))


def test_grl_conv2d_matches_unfused():
    """ops.grl_conv2d (the reversal folded into cuDNN's dgrad, DGcommon.py:73-74,106-107) against conv(grad_reverse(x))
    with the stand-alone GRL kernel: same output, input gradient to fp32 rounding, weight / bias gradients identical."""
    ops = _ops()
    g = synth.gen(13)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for cin, cout, stride, shape in [(256, 256, (2, 4), (2, 256, 38, 64)), (64, 32, 2, (1, 64, 19, 32))]:
            conv = torch.nn.Conv2d(cin, cout, 3, stride=stride).to(DEV)
            x = torch.randn(*shape, generator=g).to(DEV).requires_grad_(True)
            y1 = ops.grl_conv2d(x, conv)
            go = torch.randn(y1.shape, generator=g).to(DEV)
            y1.backward(go)
            g1, gw1, gb1 = x.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone()
            x.grad = conv.weight.grad = conv.bias.grad = None
            y2 = conv(ops.grad_reverse(x))
            y2.backward(go)
            assert torch.equal(y1, y2)
            torch.testing.assert_close(g1, x.grad, rtol=1e-5, atol=1e-6 * float(x.grad.abs().max()))
            torch.testing.assert_close(gw1, conv.weight.grad, rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(gb1, conv.bias.grad, rtol=1e-5, atol=1e-5)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------------------------- FCOS loss tail
def test_fcos_loss_tail_forward_backward():
    """ops.fcos_loss (fcos.py:149-202 fused) against the golden losses AND gradients produced by the reference's
    own FCOSHead.compute_loss + autograd on CPU (oracle/gen_golden.py::gen_fcos_loss), and against the C oracle."""
    from pathlib import Path
    from oracle import gen_golden as G
    ops = _ops()
    gold = np.load(Path(__file__).parent / "golden" / "fcos_loss.npz")
    anchors, npl, gts, labels = G.fcos_inputs()
    ho = {k: v.to(DEV).requires_grad_(True) for k, v in G.fcos_loss_inputs().items()}
    a = torch.from_numpy(anchors).to(DEV)
    idx, cls_t, box_t, onehot = ops.fcos_assign(a, [g.to(DEV) for g in gts], npl, 1.5,
                                                gt_labels=[l.to(DEV) for l in labels], num_classes=9)
    out = ops.fcos_loss(ho["cls_logits"], ho["bbox_regression"], ho["bbox_ctrness"], a, cls_t, box_t)
    ref = O.fcos_loss(ho["cls_logits"].detach().cpu().numpy(), ho["bbox_regression"].detach().cpu().numpy(),
                      ho["bbox_ctrness"].detach().cpu().numpy(), anchors, cls_t.cpu().numpy(), box_t.cpu().numpy())
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref, rtol=1e-5)
    np.testing.assert_allclose(out[:3].detach().cpu().numpy(), gold["losses"], rtol=1e-5)
    assert float(out[3]) == float((cls_t >= 0).sum())
    (out[0] + 2.0 * out[1] + 3.0 * out[2]).backward()
    for name, key in (("cls_logits", "grad_cls"), ("bbox_regression", "grad_reg"), ("bbox_ctrness", "grad_ctr")):
        g, want = ho[name].grad.cpu().numpy(), gold[key]
        np.testing.assert_allclose(g, want, rtol=1e-4, atol=1e-5 * np.abs(want).max())
    # background locations get exactly zero regression / centre-ness gradient
    bg = (cls_t < 0).cpu().numpy()
    assert not ho["bbox_regression"].grad.cpu().numpy()[bg].any() and not ho["bbox_ctrness"].grad.cpu().numpy()[bg].any()
    # no foreground at all: the three losses are finite and divided by 1 (fcos.py:197-200)
    none = torch.full_like(cls_t, -1)
    z = ops.fcos_loss(ho["cls_logits"].detach(), ho["bbox_regression"].detach(), ho["bbox_ctrness"].detach(), a, none, box_t)
    assert float(z[3]) == 0.0 and float(z[1]) == 0.0 and float(z[2]) == 0.0 and np.isfinite(float(z[0]))



# ------------------------------------------------------------------------------- input side
def test_image_batch_and_fused_transform():
    """ops.image_batch / detector.FusedTransform == GeneralizedRCNNTransform (TV transform.py:102-153): against the C
    oracle (pinned on torchvision's CPU transform) and against torchvision's own transform on the GPU, mixed sizes."""
    from torchvision.models.detection.transform import GeneralizedRCNNTransform
    from dgod_b200.detector import FusedTransform
    ops = _ops()
    g = synth.gen(21)
    imgs = [torch.rand(3, h, w, generator=g) for h, w in [(200, 333), (150, 400), (260, 180), (97, 101)]]
    boxes = [synth.random_boxes(4, im.shape[1], im.shape[2], g) for im in imgs]
    for (mn, mx, mean, std) in [(150, 300, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]), (224, 260, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])]:
        ref, sizes = O.image_batch([i.numpy() for i in imgs], mean, std, mn, mx)
        got, got_sizes = ops.image_batch([i.to(DEV) for i in imgs], mean, std, mn, mx)
        assert got_sizes == sizes and tuple(got.shape) == ref.shape
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=1e-6)
        for train in (True, False):
            tv = GeneralizedRCNNTransform(mn, mx, mean, std).to(DEV).train(train)
            mine = FusedTransform(mn, mx, mean, std).to(DEV).train(train)
            targets = [{"boxes": b.to(DEV), "labels": torch.ones(4, dtype=torch.int64, device=DEV)} for b in boxes]
            il_tv, tg_tv = tv([i.to(DEV) for i in imgs], [dict(t) for t in targets])
            il, tg = mine([i.to(DEV) for i in imgs], [dict(t) for t in targets])
            assert il.image_sizes == il_tv.image_sizes and il.tensors.shape == il_tv.tensors.shape
            torch.testing.assert_close(il.tensors, il_tv.tensors, rtol=0, atol=1e-6)
            for a, b in zip(tg, tg_tv):
                assert torch.equal(a["boxes"], b["boxes"]) and torch.equal(a["labels"], b["labels"])
            assert torch.equal(targets[0]["boxes"], boxes[0].to(DEV))            # the caller's targets are not modified
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.image_batch(imgs, [0.0] * 3, [1.0] * 3, 150, 300)


def test_image_batch_uint8_input():
    """uint8 images (0..255, as the decoder yields them): the `/ 255.0` of DrivingDataset.py:53 happens on load inside the
    kernel — bit-identical to dividing on the host and feeding floats, a quarter of the host->device bytes."""
    from dgod_b200.detector import FusedTransform
    ops = _ops()
    g = synth.gen(22)
    u8 = [torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8) for h, w in [(200, 333), (150, 400), (97, 101)]]
    as_float = [(i / 255.0).float() for i in u8]                  # what the reference's dataset hands on
    for (mn, mx, mean, std) in [(150, 300, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]), (224, 260, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])]:
        a, sa = ops.image_batch([i.to(DEV) for i in u8], mean, std, mn, mx)
        b, sb = ops.image_batch([i.to(DEV) for i in as_float], mean, std, mn, mx)
        assert sa == sb and torch.equal(a, b)
        ref, _ = O.image_batch([i.numpy() for i in as_float], mean, std, mn, mx)
        np.testing.assert_allclose(a.cpu().numpy(), ref, rtol=0, atol=1e-6)
    il, _ = FusedTransform(150, 300, [0.0] * 3, [1.0] * 3).to(DEV).eval()([i.to(DEV) for i in u8])
    il2, _ = FusedTransform(150, 300, [0.0] * 3, [1.0] * 3).to(DEV).eval()([i.to(DEV) for i in as_float])
    assert torch.equal(il.tensors, il2.tensors)
    with pytest.raises(RuntimeError, match="all float32 or all uint8"):
        ops.image_batch([u8[0].to(DEV), as_float[1].to(DEV)], [0.0] * 3, [1.0] * 3, 150, 300)


def test_image_batch_more_images_than_one_launch_takes():
    """20 images (> 16 per launch): the second launch writes behind the first; single-channel input; tiny images."""
    ops = _ops()
    g = synth.gen(33)
    imgs = [torch.rand(1, 40 + 3 * i, 55 + 2 * i, generator=g) for i in range(20)]
    ref, sizes = O.image_batch([i.numpy() for i in imgs], [0.3], [0.7], 48, 80)
    got, got_sizes = ops.image_batch([i.to(DEV) for i in imgs], [0.3], [0.7], 48, 80)
    assert got_sizes == sizes
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=1e-6)
