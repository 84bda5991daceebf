"""Drop-in parity at model level: torchvision's own Faster R-CNN / FCOS — the classes
the reference's fasterrcnn.py / fcos.py fork — run once on the stock torchvision CUDA ops and once
with `dgod_b200.patch.patch()` + `patch_model()` rebinding every call site to the sm_100a kernels.
Same weights, same inputs, same RNG: training losses must agree to RoIAlign's 1e-5 tolerance and
the eval detections must be the same set."""
import numpy as np
import pytest
import torch

from dgod_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inputs(n, h, w, seed):
    imgs = [i.to(DEV) for i in synth.random_images(n, h, w, seed)]
    targets, _ = synth.random_targets(n, 6, h, w, seed)
    return imgs, [{k: v.to(DEV) for k, v in t.items()} for t in targets]


def _match_detections(a, b, atol=2e-3):
    """fraction of detections of `a` that have a partner in `b` (same label, box within atol)"""
    if len(a["boxes"]) == 0:
        return 1.0
    hit = 0
    for box, lab in zip(a["boxes"], a["labels"]):
        cand = b["boxes"][b["labels"] == lab]
        if len(cand) and float((cand - box).abs().max(1).values.min()) <= atol:
            hit += 1
    return hit / len(a["boxes"])


@pytest.fixture
def exact_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic = old


def test_faster_rcnn_patched_equals_stock(exact_convs):
    from torchvision.models.detection import fasterrcnn_resnet50_fpn
    from dgod_b200 import patch
    torch.manual_seed(0)
    model = fasterrcnn_resnet50_fpn(weights=None, weights_backbone=None, num_classes=9, min_size=256, max_size=384,
                                    rpn_pre_nms_top_n_train=600, rpn_post_nms_top_n_train=600,
                                    rpn_pre_nms_top_n_test=400, rpn_post_nms_top_n_test=400).to(DEV)
    imgs, targets = _inputs(2, 256, 352, 1)

    def train_losses():
        model.train()
        torch.manual_seed(77)                      # the samplers' randperm
        return {k: float(v) for k, v in model(imgs, targets).items()}

    def detections():
        model.eval()
        with torch.no_grad():
            return [{k: v.cpu() for k, v in d.items()} for d in model(imgs)]

    state = {k: v.clone() for k, v in model.state_dict().items()}   # random init has live BatchNorm: the train pass moves its statistics
    ref_l, ref_d = train_losses(), detections()
    model.load_state_dict(state)
    patch.patch()
    try:
        patch.patch_model(model)
        new_l, new_d = train_losses(), detections()
    finally:
        patch.unpatch()
    assert set(ref_l) == set(new_l)
    for k in ref_l:
        np.testing.assert_allclose(new_l[k], ref_l[k], rtol=2e-4, atol=1e-6, err_msg=k)
    for a, b in zip(ref_d, new_d):
        assert abs(len(a["boxes"]) - len(b["boxes"])) <= max(2, len(a["boxes"]) // 20)
        assert _match_detections(a, b) >= 0.95 and _match_detections(b, a) >= 0.95


def test_fcos_patched_equals_stock(exact_convs):
    """The expected training losses use the REFERENCE's location->target assignment (fcos.py:510-548,
    including its `(y1-x1)*(y2-y1)` area expression — torchvision 0.26 has since changed that line),
    computed by the CPU oracle and fed to the unchanged head loss; detections compare against the
    stock CUDA ops directly."""
    import types
    from torchvision.models.detection import fcos_resnet50_fpn
    from dgod_b200 import patch
    from oracle import cpu as O
    torch.manual_seed(0)
    model = fcos_resnet50_fpn(weights=None, weights_backbone=None, num_classes=9, min_size=256, max_size=384).to(DEV)
    imgs, targets = _inputs(2, 256, 352, 2)

    def reference_compute_loss(self, targets, head_outputs, anchors, num_anchors_per_level):
        matched = []
        for a, t in zip(anchors, targets):
            idx = O.fcos_assign(a.cpu().numpy(), num_anchors_per_level[0], num_anchors_per_level[-1],
                                t["boxes"].cpu().numpy(), t["labels"].cpu().numpy(), self.center_sampling_radius)[0]
            matched.append(torch.from_numpy(idx).to(a.device))
        return self.head.compute_loss(targets, head_outputs, anchors, matched)

    stock_compute_loss = model.compute_loss
    model.compute_loss = types.MethodType(reference_compute_loss, model)

    def train_losses():
        model.train()
        return {k: float(v) for k, v in model(imgs, targets).items()}

    def detections():
        model.eval()
        model.score_thresh = 0.01                  # random-init scores sit near sigmoid(-4.6)
        with torch.no_grad():
            return [{k: v.cpu() for k, v in d.items()} for d in model(imgs)]

    state = {k: v.clone() for k, v in model.state_dict().items()}
    ref_l, ref_d = train_losses(), detections()
    model.load_state_dict(state)
    model.compute_loss = stock_compute_loss
    patch.patch()
    try:
        patch.patch_model(model)
        new_l, new_d = train_losses(), detections()
    finally:
        patch.unpatch()
    for k in ref_l:
        np.testing.assert_allclose(new_l[k], ref_l[k], rtol=1e-5, atol=1e-7, err_msg=k)   # assignment is bit-exact
    for a, b in zip(ref_d, new_d):
        assert len(a["boxes"]) == len(b["boxes"])
        assert _match_detections(a, b, 1e-4) == 1.0 and _match_detections(b, a, 1e-4) == 1.0
