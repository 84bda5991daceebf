"""Pins the C oracle against the installed torchvision CPU ops (the reference's real arithmetic,
/root/reference/requirements.txt:17).  CPU only; torchvision is present on both boxes."""
import math

import numpy as np
import pytest
import torch
import torchvision
from torchvision.models.detection import _utils as det_utils
from torchvision.ops import boxes as box_ops
from torchvision.ops import poolers

from dgod_b200 import synth
from oracle import cpu as O


def _boxes(n, seed, h=800, w=1333):
    return synth.random_boxes(n, h, w, synth.gen(seed))


# ------------------------------------------------------------------------------- box_iou
@pytest.mark.parametrize("m,n,seed", [(1, 1, 0), (20, 500, 1), (100, 2000, 2)])
def test_box_iou_bit_exact(m, n, seed):
    a, b = _boxes(m, seed), _boxes(n, seed + 100)
    ref = box_ops.box_iou(a, b).numpy()
    got = O.box_iou(a.numpy(), b.numpy())
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))


# ------------------------------------------------------------------------------- nms
@pytest.mark.parametrize("n,thr,seed", [(1, 0.5, 0), (50, 0.5, 1), (1000, 0.7, 2), (3000, 0.6, 3), (3000, 0.3, 4)])
def test_nms_bit_exact(n, thr, seed):
    g = synth.gen(seed)
    boxes = _boxes(n, seed)
    scores = torch.rand(n, generator=g)          # ties allowed: nms itself sorts stably
    ref = box_ops.nms(boxes, scores, thr).numpy()
    got = O.nms(boxes.numpy(), scores.numpy(), thr)
    assert np.array_equal(ref, got)


def test_nms_ties_and_threshold_edges():
    # equal scores: stable order (lower index first) — SURVEY.md §8c probe
    b = torch.tensor([[0, 0, 10, 10], [20, 20, 30, 30], [40, 40, 50, 50], [60, 60, 70, 70]], dtype=torch.float32)
    s = torch.tensor([0.5, 0.5, 0.5, 0.5])
    assert O.nms(b.numpy(), s.numpy(), 0.5).tolist() == box_ops.nms(b, s, 0.5).tolist() == [0, 1, 2, 3]
    s = torch.tensor([0.3, 0.9, 0.3, 0.9])
    assert O.nms(b.numpy(), s.numpy(), 0.5).tolist() == box_ops.nms(b, s, 0.5).tolist() == [1, 3, 0, 2]
    # IoU exactly at the threshold is kept (strict >): two boxes with IoU = 0.5
    b = torch.tensor([[0, 0, 2, 1], [0, 0, 1, 1]], dtype=torch.float32)
    s = torch.tensor([0.9, 0.8])
    assert O.nms(b.numpy(), s.numpy(), 0.5).tolist() == box_ops.nms(b, s, 0.5).tolist() == [0, 1]
    # double comparison: fp32 IoU 0.6000000238 vs thr 0.6 -> suppressed
    b = torch.tensor([[0, 0, 5, 1], [0, 0, 3, 1]], dtype=torch.float32)
    assert O.nms(b.numpy(), s.numpy(), 0.6).tolist() == box_ops.nms(b, s, 0.6).tolist() == [0]
    # fp32 IoU = float32(0.7) = 0.69999999 vs thr 0.7 -> both kept
    b = torch.tensor([[0, 0, 10, 1], [0, 0, 7, 1]], dtype=torch.float32)
    assert O.nms(b.numpy(), s.numpy(), 0.7).tolist() == box_ops.nms(b, s, 0.7).tolist() == [0, 1]
    # duplicates and empty
    b = torch.tensor([[0, 0, 5, 5], [0, 0, 5, 5], [0, 0, 5, 5]], dtype=torch.float32)
    s = torch.tensor([0.1, 0.2, 0.3])
    assert O.nms(b.numpy(), s.numpy(), 0.5).tolist() == box_ops.nms(b, s, 0.5).tolist() == [2]
    assert O.nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.5).tolist() == []


@pytest.mark.parametrize("n,groups,thr,seed", [(1000, 5, 0.7, 0), (1001, 5, 0.7, 1), (4000, 8, 0.5, 2), (5000, 9, 0.6, 3), (300, 3, 0.5, 4)])
def test_batched_nms_matches_torchvision(n, groups, thr, seed):
    g = synth.gen(seed)
    boxes = _boxes(n, seed, 608, 1024)
    scores = synth.distinct_scores(n, g)         # tie order is implementation-defined upstream
    idxs = torch.randint(0, groups, (n,), generator=g)
    ref = box_ops.batched_nms(boxes, scores, idxs, thr).numpy()
    got = O.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), thr)      # same mode rule
    assert np.array_equal(ref, got)
    # the two modes agree on these inputs — exercised separately
    for mode in (0, 1):
        got_m = O.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), thr, mode=mode)
        ref_m = (box_ops._batched_nms_vanilla if mode == 0 else box_ops._batched_nms_coordinate_trick)(
            boxes, scores, idxs, thr).numpy()
        assert np.array_equal(ref_m, got_m)


def test_batched_nms_tied_scores_same_set_and_score_order():
    g = synth.gen(7)
    n = 2000
    boxes = _boxes(n, 7)
    scores = (torch.randint(0, 50, (n,), generator=g).float() / 50.0)  # heavy ties
    idxs = torch.randint(0, 5, (n,), generator=g)
    ref = box_ops.batched_nms(boxes, scores, idxs, 0.7)
    got = O.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), 0.7)
    # torch's final sort is unstable (TV ops/boxes.py:120): equal scores may permute
    assert sorted(ref.tolist()) == sorted(got.tolist())
    assert np.array_equal(scores[ref].numpy(), scores.numpy()[got])
    # the oracle's tie rule: ascending index among equal scores
    s = scores.numpy()[got]
    for i in range(len(got) - 1):
        assert s[i] > s[i + 1] or (s[i] == s[i + 1] and got[i] < got[i + 1])


# ------------------------------------------------------------------------------- Matcher
@pytest.mark.parametrize("m,n,hi,lo,lq,seed", [(20, 5000, 0.7, 0.3, True, 0), (100, 2000, 0.5, 0.5, False, 1),
                                               (1, 300, 0.7, 0.3, True, 2), (7, 64, 0.5, 0.5, True, 3)])
def test_matcher_bit_exact(m, n, hi, lo, lq, seed):
    gt, pr = _boxes(m, seed), _boxes(n, seed + 50)
    q = box_ops.box_iou(gt, pr)
    ref = det_utils.Matcher(hi, lo, allow_low_quality_matches=lq)(q.clone()).numpy()
    got = O.matcher(q.numpy(), hi, lo, lq)
    assert np.array_equal(ref, got)


def test_matcher_threshold_is_fp32_and_low_quality_ties():
    q = torch.tensor([[np.float32(0.7), np.float32(0.3), 0.29999998, 0.0, 0.2, 0.2]], dtype=torch.float32)
    ref = det_utils.Matcher(0.7, 0.3, True)(q.clone()).numpy()
    got = O.matcher(q.numpy(), 0.7, 0.3, True)
    assert np.array_equal(ref, got)
    q = torch.tensor([[0.1, 0.2, 0.2, 0.0], [0.0, 0.0, 0.0, 0.0], [0.6, 0.1, 0.6, 0.0]], dtype=torch.float32)
    ref = det_utils.Matcher(0.7, 0.3, True)(q.clone()).numpy()   # all-zero row: every column ties
    got = O.matcher(q.numpy(), 0.7, 0.3, True)
    assert np.array_equal(ref, got)


def test_rpn_and_roi_assign_follow_torchvision():
    from torchvision.models.detection.rpn import RegionProposalNetwork
    from torchvision.models.detection.roi_heads import RoIHeads
    gt, anchors = _boxes(20, 3, 608, 1024), _boxes(20000, 4, 608, 1024)
    labels_gt = torch.randint(1, 9, (20,), generator=synth.gen(5))
    rpn = RegionProposalNetwork.__new__(RegionProposalNetwork)
    rpn.box_similarity = box_ops.box_iou
    rpn.proposal_matcher = det_utils.Matcher(0.7, 0.3, allow_low_quality_matches=True)
    lab, mb = RegionProposalNetwork.assign_targets_to_anchors(rpn, [anchors, anchors], [{"boxes": gt}, {"boxes": gt[:0]}])
    idx_o, lab_o, mb_o = O.rpn_assign(gt.numpy(), anchors.numpy())
    assert np.array_equal(lab[0].numpy(), lab_o) and np.array_equal(mb[0].numpy(), mb_o)
    _, lab_e, mb_e = O.rpn_assign(np.zeros((0, 4), np.float32), anchors.numpy())
    assert np.array_equal(lab[1].numpy(), lab_e) and np.array_equal(mb[1].numpy(), mb_e)
    rh = RoIHeads.__new__(RoIHeads)
    rh.proposal_matcher = det_utils.Matcher(0.5, 0.5, allow_low_quality_matches=False)
    props = torch.cat([_boxes(2000, 6, 608, 1024), gt])
    mi, lb = RoIHeads.assign_targets_to_proposals(rh, [props, props], [gt, gt[:0]], [labels_gt, labels_gt[:0]])
    ci_o, lb_o = O.roi_assign(gt.numpy(), labels_gt.numpy(), props.numpy())
    assert np.array_equal(mi[0].numpy(), ci_o) and np.array_equal(lb[0].numpy(), lb_o)
    ci_e, lb_e = O.roi_assign(np.zeros((0, 4), np.float32), np.zeros(0, np.int64), props.numpy())
    assert np.array_equal(mi[1].numpy(), ci_e) and np.array_equal(lb[1].numpy(), lb_e)


# ------------------------------------------------------------------------------- RoIAlign
def _roi_cases(H, W, scale):
    img_h, img_w = H / scale, W / scale
    rois = [
        [0, 10.3, 20.7, 200.2, 180.9], [1, 0, 0, img_w, img_h],          # generic, full image
        [0, 50, 50, 50.2, 50.3],                                        # smaller than a pixel (legacy max(.,1))
        [1, -30, -40, 60, 70],                                          # negative origin
        [0, img_w - 20, img_h - 20, img_w + 90, img_h + 80],            # past the edge (sample skip)
        [1, img_w + 40, img_h + 40, img_w + 80, img_h + 90],            # fully outside
        [0, 100, 100, 100, 100],                                        # degenerate
    ]
    return torch.tensor(rois, dtype=torch.float32)


@pytest.mark.parametrize("sr,aligned,scale", [(2, False, 0.25), (2, False, 0.125), (0, False, 0.25), (2, True, 0.25), (3, False, 0.0625)])
def test_roi_align_forward_bit_exact(sr, aligned, scale):
    H, W = 48, 64
    x = torch.randn(2, 5, H, W, generator=synth.gen(0))
    rois = torch.cat([_roi_cases(H, W, scale),
                      synth.rois_from_boxes([_boxes(40, 1, H / scale, W / scale), _boxes(40, 2, H / scale, W / scale)])])
    ref = torchvision.ops.roi_align(x, rois, (7, 7), scale, sr, aligned).numpy()
    got = O.roi_align_fwd(x.numpy(), rois.numpy(), scale, 7, 7, sr, aligned)
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))


@pytest.mark.parametrize("sr,scale", [(2, 0.25), (0, 0.125)])
def test_roi_align_backward_matches(sr, scale):
    H, W = 40, 56
    rois = torch.cat([_roi_cases(H, W, scale), synth.rois_from_boxes([_boxes(30, 3, H / scale, W / scale), _boxes(30, 4, H / scale, W / scale)])])
    x = torch.randn(2, 4, H, W, generator=synth.gen(1), requires_grad=True)
    out = torchvision.ops.roi_align(x, rois, (7, 7), scale, sr, False)
    go = torch.randn(out.shape, generator=synth.gen(2))
    out.backward(go)
    got = O.roi_align_bwd(go.numpy(), tuple(x.shape), rois.numpy(), scale, sr)
    # same kernel, same sequential accumulation order on CPU -> bit-exact
    assert np.array_equal(x.grad.numpy().view(np.uint32), got.view(np.uint32))


def test_level_mapper_incl_exact_boundaries():
    sides = [111.99, 112.0, 112.01, 223.99, 224.0, 224.01, 447.9, 448.0, 448.1, 20.0, 900.0, 56.0, 1.0]
    boxes = torch.tensor([[10.0, 20.0, 10.0 + s, 20.0 + s] for s in sides] + [[0, 0, 100, 501.76], [5, 5, 5, 5]], dtype=torch.float32)
    boxes = torch.cat([boxes, _boxes(2000, 11)])
    ref = poolers.LevelMapper(2, 5)([boxes]).numpy()
    got = O.level_map(boxes.numpy(), 2, 5)
    assert np.array_equal(ref, got)


def test_multiscale_roi_align_matches_pooler():
    img_h, img_w = 200, 264
    feats = synth.random_features(2, 6, img_h, img_w, seed=3)
    boxes = [_boxes(60, 5, img_h, img_w), _boxes(50, 6, img_h, img_w)]
    pool = torchvision.ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2)
    x = {str(i): f.clone().requires_grad_(True) for i, f in enumerate(feats)}
    ref = pool(x, boxes, [(img_h, img_w)] * 2)
    rois = synth.rois_from_boxes(boxes)
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
    assert pool.scales == scales
    got = O.msroi_align_fwd([f.numpy() for f in feats], rois.numpy(), scales, 7, 7, 2, 2, 5)
    assert np.array_equal(ref.detach().numpy().view(np.uint32), got.view(np.uint32))
    go = torch.randn(ref.shape, generator=synth.gen(9))
    ref.backward(go)
    grads = O.msroi_align_bwd(go.numpy(), [tuple(f.shape) for f in feats], rois.numpy(), scales, 2, 2, 5)
    for i, gr in enumerate(grads):
        # the pooler scatters per level (different accumulation order across levels only)
        np.testing.assert_allclose(x[str(i)].grad.numpy(), gr, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------- anchors / decode / RPN filter
def test_anchors_and_decode():
    from torchvision.models.detection.anchor_utils import AnchorGenerator
    from torchvision.models.detection.image_list import ImageList
    ag = AnchorGenerator(((32,), (64,), (128,), (256,), (512,)), ((0.5, 1.0, 2.0),) * 5)
    H, W = 608, 1024
    grids = [(152, 256), (76, 128), (38, 64), (19, 32), (10, 16)]
    feats = [torch.zeros(1, 1, h, w) for h, w in grids]
    ref = ag(ImageList(torch.zeros(1, 3, H, W), [(600, 999)]), feats)[0]
    got = np.concatenate([O.grid_anchors(ag.cell_anchors[l].numpy(), h, w, H // h, W // w) for l, (h, w) in enumerate(grids)])
    assert np.array_equal(ref.numpy(), got)
    g = synth.gen(0)
    n = 5000
    sel = torch.randint(0, len(ref), (n,), generator=g)
    rel = torch.randn(n, 4, generator=g) * 0.5
    rel[0, 2] = 10.0  # hits the log(1000/16) clamp
    coder = det_utils.BoxCoder((1.0, 1.0, 1.0, 1.0))
    dref = coder.decode_single(rel, ref[sel]).numpy()
    dgot = O.box_decode(rel.numpy(), ref[sel].numpy())
    np.testing.assert_allclose(dref, dgot, rtol=2e-6, atol=1e-4)
    coder2 = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
    rel9 = torch.randn(n, 36, generator=g)
    np.testing.assert_allclose(coder2.decode_single(rel9, ref[sel]).numpy(),
                               O.box_decode(rel9.numpy(), ref[sel].numpy(), (10.0, 10.0, 5.0, 5.0)), rtol=2e-6, atol=1e-4)


def test_rpn_filter_matches_torchvision():
    from torchvision.models.detection.rpn import RegionProposalNetwork
    g = synth.gen(4)
    npl = [3 * 38 * 64, 3 * 19 * 32, 3 * 10 * 16, 3 * 5 * 8, 3 * 3 * 4]
    A = sum(npl)
    img = (300.0, 500.0)
    props = torch.stack([synth.random_boxes(A, 330, 540, g), synth.random_boxes(A, 330, 540, g)]) - 15.0
    obj = torch.randn(2, A, generator=g)
    rpn = RegionProposalNetwork.__new__(RegionProposalNetwork)
    torch.nn.Module.__init__(rpn)
    rpn._pre_nms_top_n = {"training": 600, "testing": 300}
    rpn._post_nms_top_n = {"training": 500, "testing": 300}
    rpn.nms_thresh, rpn.score_thresh, rpn.min_size = 0.7, 0.0, 1e-3
    rpn.train()
    fb, fs = RegionProposalNetwork.filter_proposals(rpn, props, obj.reshape(-1, 1), [img, img], npl)
    for i in range(2):
        ob, os_ = O.rpn_filter_image(props[i].numpy(), obj[i].numpy(), npl, 600, 500, 0.7, 1e-3, 0.0, *img)
        assert len(ob) == len(fb[i])
        assert np.array_equal(fb[i].numpy(), ob)                       # boxes bit-exact, same order
        np.testing.assert_allclose(fs[i].numpy(), os_, rtol=2e-7, atol=0)  # sigmoid within 1 ulp


def test_detect_candidates_and_grl():
    g = synth.gen(8)
    n, ncls = 512, 9
    props = synth.random_boxes(n, 600, 1000, g)
    logits, reg = torch.randn(n, ncls, generator=g) * 2, torch.randn(n, ncls * 4, generator=g) * 0.5
    coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
    pb = box_ops.clip_boxes_to_image(coder.decode(reg, [props]), (600, 1000))[:, 1:]
    ps = torch.softmax(logits, -1)[:, 1:]
    cb, cs, cl, cv = O.detect_candidates(logits.numpy(), reg.numpy(), props.numpy(), 600, 1000)
    np.testing.assert_allclose(pb.numpy(), cb, rtol=2e-6, atol=2e-4)
    np.testing.assert_allclose(ps.numpy(), cs, rtol=1e-6, atol=1e-9)
    assert np.array_equal(cl, np.tile(np.arange(1, ncls), (n, 1)))
    ws, hs = cb[..., 2] - cb[..., 0], cb[..., 3] - cb[..., 1]
    assert np.array_equal(cv.astype(bool), (cs > np.float32(0.05)) & (ws >= np.float32(1e-2)) & (hs >= np.float32(1e-2)))
    x = torch.randn(1000, generator=g)
    assert np.array_equal((x.neg() * 0.1).numpy(), O.grl_scale(x.numpy(), 0.1))


def test_image_batch_matches_torchvision_transform():
    """oracle o_image_resize_pad == GeneralizedRCNNTransform.forward (TV transform.py:102-153, the transform
    fasterrcnn.py:439-441 / fcos.py:483 construct): mixed image sizes, both scale regimes, non-trivial mean/std."""
    from torchvision.models.detection.transform import GeneralizedRCNNTransform
    g = synth.gen(21)
    imgs = [torch.rand(3, h, w, generator=g) for h, w in [(200, 333), (150, 400), (260, 180), (97, 101)]]
    boxes = [synth.random_boxes(4, im.shape[1], im.shape[2], g) for im in imgs]
    for (mn, mx, mean, std) in [(150, 300, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]), (224, 260, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])]:
        tr = GeneralizedRCNNTransform(mn, mx, mean, std).eval()
        il, tg = tr([i.clone() for i in imgs], [{"boxes": b.clone()} for b in boxes])
        out, sizes = O.image_batch([i.numpy() for i in imgs], mean, std, mn, mx)
        assert [tuple(s) for s in il.image_sizes] == sizes and tuple(il.tensors.shape) == out.shape
        np.testing.assert_allclose(out, il.tensors.numpy(), rtol=0, atol=1e-6)     # 1-2 ulp: the blend's own rounding
        for t, b, im, (oh, ow) in zip(tg, boxes, imgs, sizes):                       # resize_boxes, transform.py:305-316
            r = torch.tensor([ow / im.shape[2], oh / im.shape[1]] * 2)
            torch.testing.assert_close(t["boxes"], b * r, rtol=1e-6, atol=1e-4)
