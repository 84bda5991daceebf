"""The benchmarked Faster R-CNN path (dgod_b200.detector / dgod_b200.dg, the objects bench.py times) on the
GPU against golden vectors produced by the REFERENCE's own modules on the CPU (oracle/gen_golden.py ->
tests/golden/frcnn_hotpath.npz: fasterrcnn.RegionProposalNetworkWILDS / RoIHeadsWILDS, fasterrcnn.py:90-305).

The reference's samplers draw from torch's RNG; the fixture records what they picked and the mirror is given the
same selection through `sampler_keys`, so proposals, anchor labels, the sampled `box_labels` that DGFRCNN's hook
hands to the DG heads (DGFRCNN.py:89-91) and the per-image losses can be compared value for value."""
from pathlib import Path

import numpy as np
import pytest
import torch

from dgod_b200 import synth
from oracle import gen_golden as G

GOLD = Path(__file__).parent / "golden"


@pytest.fixture
def exact_convs():
    """fp32 convolutions / matmuls without TF32 so that the comparison with the CPU run is at fp32 rounding."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _keys(pos_bits, neg_bits, n):
    """Sampler keys that force detector.BalancedSampler to the recorded selection: 0 for picked entries, 1 for the rest."""
    pos = np.unpackbits(pos_bits, axis=1)[:, :n].astype(bool)
    neg = np.unpackbits(neg_bits, axis=1)[:, :n].astype(bool)
    return torch.from_numpy(np.where(pos | neg, 0.0, 1.0).astype(np.float32)), pos, neg


def _modules():
    from dgod_b200.detector import RegionProposalNetwork, RoIHeads
    rpn = RegionProposalNetwork(256, ((32,), (64,), (128,), (256,), (512,)), ((0.5, 1.0, 2.0),) * 5,
                                pre_nms_top_n=dict(training=600, testing=600), post_nms_top_n=dict(training=600, testing=600))
    roi = RoIHeads(256, 9)
    G.seeded_module_weights(rpn, 1)        # the parameter names equal the reference's, so the draws are identical
    G.seeded_module_weights(roi, 2)
    return rpn.cuda().train(), roi.cuda().train()


@pytest.mark.gpu
@pytest.mark.parametrize("channels_last", [False, True])
def test_frcnn_rpn_and_roi_heads_match_reference_golden(exact_convs, channels_last):
    gold = np.load(GOLD / "frcnn_hotpath.npz")
    h, w = G.HOT["img"]
    B = G.HOT["batch"]
    features, targets = G.hotpath_inputs()
    feats = {k: v.cuda() for k, v in features.items()}
    if channels_last:
        feats = {k: v.contiguous(memory_format=torch.channels_last) for k, v in feats.items()}
    tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
    rpn, roi = _modules()

    # ---- RPN: proposals, anchor labels, per-image losses (fasterrcnn.py:142-196)
    n_anchor = gold["anchor_labels"].shape[1]
    keys, pos, neg = _keys(gold["rpn_pos"], gold["rpn_neg"], n_anchor)
    (pb, ps, pc), rpn_losses = rpn((B, 3, h, w), [(h, w)] * B, feats, tg, keys.cuda())
    counts = pc.tolist()
    for i in range(B):
        got, want = pb[i, :counts[i]].cpu().numpy(), gold["proposals"][i]
        assert got.shape == want.shape
        # the same SET of boxes; the order is only defined up to scores that are equal or 1 ulp apart (the head's
        # convolutions round differently on the two devices, and TV's final sort is unstable)
        key = lambda a: a[np.lexsort(np.round(a, 2).T[::-1])]
        np.testing.assert_allclose(key(got), key(want), rtol=0, atol=2e-4)
        assert (np.abs(got - want).max(1) > 2e-4).mean() < 0.02
    assert np.array_equal(rpn.last_anchor_labels.cpu().numpy().astype(np.int8), gold["anchor_labels"])   # bit-exact
    lab = gold["anchor_labels"]
    assert (lab[pos] == 1).all() and (lab[neg] == 0).all()                  # the fixture's picks are what the sampler may pick
    for k in ("loss_objectness", "loss_rpn_box_reg"):
        np.testing.assert_allclose(rpn_losses[k].detach().cpu().numpy(), gold[k], rtol=1e-5, atol=1e-7)

    # ---- RoI heads on the reference's proposals: sampled labels (the DG hand-off), pooled features, losses
    n_prop = gold["proposals"].shape[1] + G.HOT["n_gt"]
    keys, pos, neg = _keys(gold["roi_pos"], gold["roi_neg"], n_prop)
    grabbed = {}
    roi.box_head.register_forward_hook(lambda m, i, o: grabbed.update(pooled=i[0], labels=i[1]))
    props = [torch.from_numpy(gold["proposals"][i]).cuda() for i in range(B)]
    det, roi_losses, box_features, box_labels = roi(feats, props, [(h, w)] * B, tg, keys.cuda())
    got_labels = torch.stack(box_labels).cpu().numpy()
    assert got_labels.dtype == np.int64 and np.array_equal(got_labels, gold["roi_labels"])              # bit-exact
    assert torch.equal(torch.stack(grabbed["labels"]), torch.stack(box_labels))                         # what the hook sees
    assert box_features.shape == (B * 512, 1024)
    np.testing.assert_allclose(grabbed["pooled"].double().sum(dim=(1, 2, 3)).cpu().numpy(), gold["pooled_sum"],
                               rtol=1e-5, atol=1e-5)
    for k in ("loss_classifier", "loss_box_reg"):
        np.testing.assert_allclose(roi_losses[k].detach().cpu().numpy(), gold[k], rtol=1e-5, atol=1e-7)
    # gradients flow through the RoIAlign backward into every FPN level that received RoIs
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in feats.items()}
    _, l2, _, _ = roi(leaf, props, [(h, w)] * B, tg, keys.cuda())
    (l2["loss_classifier"].sum() + l2["loss_box_reg"].sum()).backward()
    assert all(torch.isfinite(leaf[k].grad).all() for k in "0123") and float(leaf["0"].grad.abs().sum()) > 0


@pytest.mark.gpu
def test_roi_heads_with_fewer_candidates_than_slots():
    """An image with fewer than 512 proposals no longer raises (TV roi_heads.py:615-622 returns what it has): the
    [B, 512] shape of the DG hand-off is kept, empty slots are labelled -100 and leave the losses."""
    _, roi = _modules()
    h, w = 256, 320
    feats = {k: v.cuda() for k, v in zip("0123", synth.random_features(2, 256, h, w, 3))}
    targets, _ = synth.random_targets(2, 4, h, w, 5)
    tg = [{k: v.cuda() for k, v in t.items()} for t in targets]
    props = [synth.random_boxes(n, h, w, synth.gen(40 + n)).cuda() for n in (700, 90)]
    _, losses, box_features, labels = roi(feats, props, [(h, w)] * 2, tg)
    assert box_features.shape == (2 * 512, 1024) and labels[0].shape == labels[1].shape == (512,)
    assert int((labels[0] >= 0).sum()) == 512
    assert int((labels[1] >= 0).sum()) == 94 and int((labels[1] == -100).sum()) == 512 - 94       # 90 proposals + 4 GT
    assert all(torch.isfinite(v).all() for v in losses.values())


@pytest.mark.gpu
def test_dgfrcnn_cycle_on_gpu():
    """BASELINE configs[1] at B=2: one full 8-step mode cycle of dg.DGFRCNN (the object bench.py times): finite
    losses, the DG hand-off contract (DGFRCNN.py:89-91,151-156) and gradients where DGFRCNN.py puts them."""
    from dgod_b200 import dg
    torch.manual_seed(0)
    B, D = 2, 2
    m = dg.DGFRCNN(9, B, "dg", [0.5, 0.5, 0.5, 0.05, 0.0001], num_domains=D).cuda().train()
    assert m.detector.roi_heads.box_predictor.cls_score.out_features == 10       # fasterrcnn.py:327: num_classes + 1
    assert not m.detector.backbone.body.layer1[0].conv1.weight.requires_grad      # trainable_backbone_layers=3
    opt = m.configure_optimizer(lr=1e-5)
    imgs = [i.cuda() for i in synth.random_images(B, 608, 1024, 3)]
    targets, dom = synth.random_targets(B, 6, 608, 1024, 3, n_domains=D)
    batch = (imgs, [t["boxes"].cuda() for t in targets], [t["labels"].cuda() for t in targets], dom.cuda())
    cls_w = m.detector.roi_heads.box_predictor.cls_score.weight
    for step, mode in enumerate([0, 1, 0, 2, 0, 3, 0, 4]):
        assert m.mode == mode
        loss = m.training_step(batch)
        assert torch.isfinite(loss), (step, mode)
        labels = torch.stack(m.box_labels)
        assert labels.shape == (B, 512) and labels.dtype == torch.int64
        assert int(labels.max()) <= 8 and bool(((labels >= 0) | (labels == -100)).all())
        assert m.box_features.shape == (B * 512, 1024) and m.base_feat["0"].shape[1] == 256
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if mode == 0:
            assert cls_w.grad is not None and float(cls_w.grad.abs().sum()) > 0
        if mode == 1:
            assert m.ImageDA.Conv1.weight.grad is not None and m.InsDA.dc_ip1.weight.grad is not None
            assert m.detector.backbone.fpn.inner_blocks[0][0].weight.grad is not None      # through the GRL
        if mode == 2:
            assert cls_w.grad is None and m.InsCls[0].dc_ip1.weight.grad is not None       # detector under no_grad
        if mode == 3:
            assert m.InsClsPrime[0].dc_ip1.weight.grad is not None
        opt.step()
    assert m.mode == 0 and m.sub_mode == 0
