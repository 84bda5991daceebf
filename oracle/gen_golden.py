"""Generates tests/golden/*.npz by running the REFERENCE's own modules (imported unchanged from
/root/reference) on seeded synthetic inputs.  Run here (the reference is not on the GPU box):

    python -m oracle.gen_golden

Fixtures:
  fcos_assign.npz     fcos.FCOS.compute_loss (fcos.py:503-550) matched indices and
                      fcos.FCOSHead.compute_loss (fcos.py:124-202) `gt_classes` one-hot
  frcnn_hotpath.npz   fasterrcnn.RegionProposalNetworkWILDS / RoIHeadsWILDS (fasterrcnn.py:90-305) on
                      synthetic FPN features: proposals, anchor labels, sampled RoI labels, pooled
                      features checksum, per-image losses
  fcos_post.npz       fcos.FCOS.postprocess_detections (fcos.py:552-619) on seeded head outputs: eval detections
  fcos_loss.npz       fcos.FCOSHead.compute_loss (fcos.py:124-202) on seeded head outputs: losses and gradients
  fcos_step.npz       losses + gt_classes of one training forward of fcos.fcos_resnet50_fpn with name-seeded
                      weights (the dgod_b200.dg_fcos mirror must reproduce them on the GPU)
  frcnn_step.npz      per-image losses of a full fasterrcnn.FastWILDS train step vs the restated
                      oracle.ref_dgfrcnn.RefFasterRCNN with identical weights and RNG
"""
from __future__ import annotations

import sys
import types
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))

from dgod_b200 import synth  # noqa: E402


def import_reference():
    if not REF.exists():
        raise SystemExit("/root/reference is not available on this machine")
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    warnings.filterwarnings("ignore")
    import fasterrcnn  # noqa
    import fcos  # noqa
    return fasterrcnn, fcos


# ----------------------------------------------------------------------------------------- FCOS
def fcos_inputs():
    """Anchors of a 256x320 image on 5 levels (anchor size 8*stride, fcos.py:467-469) and GT lists
    covering 0, 1, 2 and many boxes, with one pair of equal quirky areas."""
    from oracle import cpu as O
    anchors, npl = [], []
    for s in (8, 16, 32, 64, 128):
        gh, gw = -(-256 // s), -(-320 // s)
        anchors.append(O.grid_anchors(np.array([[-4 * s, -4 * s, 4 * s, 4 * s]], np.float32), gh, gw, s, s))
        npl.append(gh * gw)
    anchors = np.concatenate(anchors)
    gts, labels = [], []
    for i, m in enumerate([12, 0, 1, 2, 40]):
        g = synth.gen(100 + i)
        b = synth.random_boxes(m, 256, 320, g, log_size=(2.0, 3.2))
        if m >= 2:
            b[1] = b[0] + torch.tensor([3.0, 0.0, 3.0, 0.0])
        if m == 1:
            b[0] = torch.tensor([60.0, 50.0, 160.0, 150.0])   # certainly matched: shows the `<= 1` rule
        gts.append(b)
        labels.append(torch.randint(1, 9, (m,), generator=g))
    return anchors, npl, gts, labels


def gen_fcos(fcos):
    anchors, npl, gts, labels = fcos_inputs()
    B, N = len(gts), len(anchors)
    a = torch.from_numpy(anchors)
    targets = [{"boxes": g, "labels": l} for g, l in zip(gts, labels)]
    stub = types.SimpleNamespace(center_sampling_radius=1.5,
                                 head=types.SimpleNamespace(compute_loss=lambda t, h, an, m: m))
    matched = fcos.FCOS.compute_loss(stub, targets, None, [a] * B, npl)
    head = types.SimpleNamespace(box_coder=fcos.BoxLinearCoder(normalize_by_size=True))
    ho = {"cls_logits": torch.zeros(B, N, 9), "bbox_regression": torch.zeros(B, N, 4), "bbox_ctrness": torch.zeros(B, N, 1)}
    loss = fcos.FCOSHead.compute_loss(head, targets, ho, [a] * B, [m.clone() for m in matched])
    np.savez_compressed(OUT / "fcos_assign.npz", matched=torch.stack(matched).numpy(),
                        gt_classes=loss["gt_classes"].numpy().astype(np.uint8))
    print("fcos_assign.npz", torch.stack(matched).shape, int((torch.stack(matched) >= 0).sum()), "matched")


def fcos_loss_inputs():
    """Seeded head outputs for the inputs of fcos_inputs(): logits ~ N(0, 2), regression in (0.05, 2.05) (the head
    ends in ReLU, fcos.py:299), centre-ness logits ~ N(0, 1)."""
    anchors, npl, gts, labels = fcos_inputs()
    B, N = len(gts), len(anchors)
    g = synth.gen(4242)
    return {"cls_logits": torch.randn(B, N, 9, generator=g) * 2.0,
            "bbox_regression": torch.rand(B, N, 4, generator=g) * 2.0 + 0.05,
            "bbox_ctrness": torch.randn(B, N, 1, generator=g)}


def gen_fcos_loss(fcos):
    """fcos.FCOSHead.compute_loss (fcos.py:124-202) on seeded head outputs: the three losses and their gradients."""
    anchors, npl, gts, labels = fcos_inputs()
    B = len(gts)
    a = torch.from_numpy(anchors)
    targets = [{"boxes": g, "labels": l} for g, l in zip(gts, labels)]
    stub = types.SimpleNamespace(center_sampling_radius=1.5,
                                 head=types.SimpleNamespace(compute_loss=lambda t, h, an, m: m))
    matched = fcos.FCOS.compute_loss(stub, targets, None, [a] * B, npl)
    head = types.SimpleNamespace(box_coder=fcos.BoxLinearCoder(normalize_by_size=True))
    ho = {k: v.requires_grad_(True) for k, v in fcos_loss_inputs().items()}
    loss = fcos.FCOSHead.compute_loss(head, targets, ho, [a] * B, [m.clone() for m in matched])
    total = loss["classification"] + 2.0 * loss["bbox_regression"] + 3.0 * loss["bbox_ctrness"]
    total.backward()
    np.savez_compressed(OUT / "fcos_loss.npz",
                        losses=np.array([float(loss[k]) for k in ("classification", "bbox_regression", "bbox_ctrness")], np.float32),
                        grad_cls=ho["cls_logits"].grad.numpy(), grad_reg=ho["bbox_regression"].grad.numpy(),
                        grad_ctr=ho["bbox_ctrness"].grad.numpy())
    print("fcos_loss.npz", {k: float(loss[k]) for k in ("classification", "bbox_regression", "bbox_ctrness")})


def fcos_post_inputs():
    """Seeded head outputs for the eval post-processing on the anchors of fcos_inputs(): 3 images, class logits around the
    0.2 score threshold (so that the threshold, the top-k cut and the NMS all bite), positive box regressions."""
    anchors, npl, _, _ = fcos_inputs()
    B, N = 3, len(anchors)
    g = synth.gen(777)
    return {"cls_logits": torch.randn(B, N, 9, generator=g) * 2.5 - 2.0,
            "bbox_regression": torch.rand(B, N, 4, generator=g) * 3.0 + 0.1,
            "bbox_ctrness": torch.randn(B, N, 1, generator=g) * 1.5}, [(256, 320), (250, 300), (256, 320)]


def gen_fcos_post(fcos):
    """fcos.FCOS.postprocess_detections (fcos.py:552-619) on seeded head outputs, topk_candidates lowered to 300 so that
    the top-k cut is exercised: detections per image."""
    anchors, npl, _, _ = fcos_inputs()
    ho, shapes = fcos_post_inputs()
    a = torch.from_numpy(anchors)
    stub = types.SimpleNamespace(score_thresh=0.2, nms_thresh=0.6, detections_per_img=100, topk_candidates=300,
                                 box_coder=fcos.BoxLinearCoder(normalize_by_size=True))
    split = {k: list(v.split(npl, dim=1)) for k, v in ho.items()}
    det = fcos.FCOS.postprocess_detections(stub, split, [list(a.split(npl))] * len(shapes), shapes)
    out = {}
    for i, d in enumerate(det):
        out[f"boxes{i}"], out[f"scores{i}"], out[f"labels{i}"] = d["boxes"].numpy(), d["scores"].numpy(), d["labels"].numpy()
    np.savez_compressed(OUT / "fcos_post.npz", **out)
    print("fcos_post.npz", [len(d["boxes"]) for d in det])


# ----------------------------------------------------------------------------------------- Faster R-CNN hot path
HOT = dict(img=(256, 320), batch=2, channels=256, n_gt=6, seed=7)


def hotpath_inputs():
    h, w = HOT["img"]
    feats = synth.random_features(HOT["batch"], HOT["channels"], h, w, HOT["seed"], strides=(4, 8, 16, 32, 64))
    features = {k: f * 0.1 for k, f in zip(["0", "1", "2", "3", "pool"], feats)}
    targets, _ = synth.random_targets(HOT["batch"], HOT["n_gt"], h, w, HOT["seed"])
    return features, targets


def seeded_module_weights(mod: torch.nn.Module, seed: int):
    g = synth.gen(seed)
    with torch.no_grad():
        for _, p in sorted(mod.named_parameters()):
            p.copy_(torch.randn(p.shape, generator=g) * (0.02 if p.dim() > 1 else 0.01))


def gen_hotpath(fasterrcnn):
    from torchvision.models.detection.image_list import ImageList
    h, w = HOT["img"]
    features, targets = hotpath_inputs()
    model = fasterrcnn.FastWILDS(types.SimpleNamespace(out_channels=256), num_classes=9,
                                 rpn_pre_nms_top_n_train=600, rpn_post_nms_top_n_train=600)
    seeded_module_weights(model.rpn, 1)
    seeded_module_weights(model.roi_heads, 2)
    model.train()
    images = ImageList(torch.zeros(HOT["batch"], 3, h, w), [(h, w)] * HOT["batch"])
    torch.manual_seed(1234)
    # the samplers draw from torch's global RNG: record what they picked so that the GPU mirror can be given the
    # same selection (detector.BalancedSampler `keys`) and everything downstream compared value for value
    picks = {"rpn": [], "roi": []}

    def recording(sampler, key):
        call = sampler.__call__

        def wrapped(matched_idxs):
            pos, neg = call(matched_idxs)
            picks[key].append((torch.stack(pos), torch.stack(neg)))
            return pos, neg
        return wrapped

    model.rpn.fg_bg_sampler = recording(model.rpn.fg_bg_sampler, "rpn")
    model.roi_heads.fg_bg_sampler = recording(model.roi_heads.fg_bg_sampler, "roi")
    proposals, rpn_losses = model.rpn(images, features, targets)
    anchors = model.rpn.anchor_generator(images, list(features.values()))
    labels, _ = model.rpn.assign_targets_to_anchors(anchors, targets)
    grabbed = {}
    model.roi_heads.box_head.register_forward_hook(lambda m, i, o: grabbed.update(labels=i[1], pooled=i[0]))
    det, roi_losses = model.roi_heads(features, [p.detach() for p in proposals], images.image_sizes, targets)
    np.savez_compressed(
        OUT / "frcnn_hotpath.npz",
        proposals=np.stack([p.detach().numpy() for p in proposals]),
        anchor_labels=torch.stack(labels).numpy().astype(np.int8),
        roi_labels=torch.stack(grabbed["labels"]).numpy(),
        pooled_sum=grabbed["pooled"].detach().double().sum(dim=(1, 2, 3)).numpy(),
        rpn_pos=np.packbits(torch.cat([p for p, _ in picks["rpn"]]).numpy().astype(np.uint8), axis=1),
        rpn_neg=np.packbits(torch.cat([n for _, n in picks["rpn"]]).numpy().astype(np.uint8), axis=1),
        roi_pos=np.packbits(torch.cat([p for p, _ in picks["roi"]]).numpy().astype(np.uint8), axis=1),
        roi_neg=np.packbits(torch.cat([n for _, n in picks["roi"]]).numpy().astype(np.uint8), axis=1),
        loss_objectness=rpn_losses["loss_objectness"].detach().numpy(),
        loss_rpn_box_reg=rpn_losses["loss_rpn_box_reg"].detach().numpy(),
        loss_classifier=roi_losses["loss_classifier"].detach().numpy(),
        loss_box_reg=roi_losses["loss_box_reg"].detach().numpy(),
        det_counts=np.array([len(d["boxes"]) for d in det]))
    print("frcnn_hotpath.npz", [tuple(p.shape) for p in proposals], {k: v.tolist() for k, v in rpn_losses.items()})


# ----------------------------------------------------------------------------------------- full step
STEP = dict(img=(224, 288), batch=2, n_gt=5, seed=11, min_size=224, max_size=320)


def gen_step(fasterrcnn):
    from oracle.ref_dgfrcnn import RefFasterRCNN
    torch.manual_seed(0)
    ref = fasterrcnn.fasterrcnn_resnet50_fpn(num_classes=8, pretrained=False, pretrained_backbone=False,
                                             min_size=STEP["min_size"], max_size=STEP["max_size"])
    mine = RefFasterRCNN(9, STEP["min_size"], STEP["max_size"])
    missing = mine.load_state_dict(ref.state_dict(), strict=True)
    ref.train(), mine.train()
    imgs = synth.random_images(STEP["batch"], *STEP["img"], STEP["seed"])
    targets, _ = synth.random_targets(STEP["batch"], STEP["n_gt"], *STEP["img"], STEP["seed"])
    out = {}
    for name, model in (("ref", ref), ("mine", mine)):
        torch.manual_seed(4321)
        det = model([i.clone() for i in imgs], [{k: v.clone() for k, v in t.items()} for t in targets])
        out[name] = {k: torch.stack([d["losses"][k] for d in det]).detach().numpy() for k in det[0]["losses"]}
    for k in out["ref"]:
        assert np.array_equal(out["ref"][k], out["mine"][k]), (k, out["ref"][k], out["mine"][k])
    np.savez_compressed(OUT / "frcnn_step.npz", **out["ref"])
    print("frcnn_step.npz restated step == reference step bit-for-bit:", {k: v.tolist() for k, v in out["ref"].items()}, missing)


# ----------------------------------------------------------------------------------------- FCOS step
FSTEP = dict(img=(224, 288), batch=3, n_gt=(6, 1, 4), seed=21, min_size=224, max_size=320)


def name_seeded_weights(mod: torch.nn.Module, seed: int):
    """Deterministic weights keyed by parameter NAME (He-scaled), so that the reference model here and
    the dgod_b200 mirror on the GPU box hold identical parameters without shipping a checkpoint."""
    g = synth.gen(seed)
    with torch.no_grad():
        for _, p in sorted(mod.named_parameters()):
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g) * (2.0 / p[0].numel()) ** 0.5)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.01)


def fcos_step_inputs():
    imgs = synth.random_images(FSTEP["batch"], *FSTEP["img"], FSTEP["seed"])
    targets = []
    for i, m in enumerate(FSTEP["n_gt"]):
        g = synth.gen(FSTEP["seed"] * 100 + i)
        targets.append({"boxes": synth.random_boxes(m, *FSTEP["img"], g, log_size=(2.5, 2.5)),
                        "labels": torch.randint(1, 9, (m,), generator=g, dtype=torch.int64)})
    return imgs, targets


def gen_fcos_step(fcos):
    """One training forward of the reference's fcos.fcos_resnet50_fpn (fcos.py:702-788) on CPU: the three
    losses and the `gt_classes` one-hot it hands to DGFCOS (fcos.py:196-202), incl. a 1-GT image."""
    model = fcos.fcos_resnet50_fpn(num_classes=9, pretrained_backbone=False, trainable_backbone_layers=3,
                                   min_size=FSTEP["min_size"], max_size=FSTEP["max_size"])
    name_seeded_weights(model, 5)
    model.train()
    imgs, targets = fcos_step_inputs()
    out = model(imgs, targets)
    gtc = out["gt_classes"].detach().numpy()
    np.savez_compressed(OUT / "fcos_step.npz",
                        classification=out["classification"].detach().numpy(),
                        bbox_regression=out["bbox_regression"].detach().numpy(),
                        bbox_ctrness=out["bbox_ctrness"].detach().numpy(),
                        gt_classes=np.packbits(gtc.astype(np.uint8), axis=None), gt_shape=np.array(gtc.shape),
                        param_abs_sum=np.array(sum(float(p.detach().double().abs().sum()) for p in model.parameters())))
    print("fcos_step.npz", {k: float(out[k]) for k in ("classification", "bbox_regression", "bbox_ctrness")},
          gtc.shape, int(gtc.sum()))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    fasterrcnn, fcos = import_reference()
    gen_fcos(fcos)
    gen_fcos_loss(fcos)
    gen_fcos_post(fcos)
    gen_fcos_step(fcos)
    gen_hotpath(fasterrcnn)
    gen_step(fasterrcnn)


if __name__ == "__main__":
    main()
