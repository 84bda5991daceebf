"""The FCOS / DGFCOS host mirror (dgod_b200/dg_fcos.py, BASELINE configs[2]) against golden vectors
produced by the reference's own fcos.py (oracle/gen_golden.py::gen_fcos_step) and its mode logic
(DGFCOS.py:153-243)."""
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import gen_golden as G

GOLD = Path(__file__).parent / "golden"


def _mirror():
    from dgod_b200 import dg_fcos
    model = dg_fcos.fcos_resnet50_fpn(num_classes=9, trainable_backbone_layers=3,
                                      min_size=G.FSTEP["min_size"], max_size=G.FSTEP["max_size"])
    G.name_seeded_weights(model, 5)
    return model


def test_fcos_mirror_holds_the_reference_parameters():
    """Same parameter names and shapes as fcos.fcos_resnet50_fpn: the name-seeded weights sum to the
    value recorded from the reference model."""
    gold = np.load(GOLD / "fcos_step.npz")
    model = _mirror()
    s = sum(float(p.detach().double().abs().sum()) for p in model.parameters())
    assert abs(s - float(gold["param_abs_sum"])) <= 1e-9 * float(gold["param_abs_sum"])
    frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
    assert any(n.startswith("backbone.body.conv1") for n in frozen) and not any("layer2" in n for n in frozen)


def test_dgfcos_mode_schedule_and_losses(monkeypatch):
    """DGFCOS.training_step walks 0,1,0,2,0,3,0,4 (DGFCOS.py:166-243); modes 2-4 call the detector once per
    image; the odd cross_entropy of [1,N,C] scores against the [1,N,C] one-hot is reproduced as written."""
    from dgod_b200 import dg_fcos

    class FakeFCOS(torch.nn.Module):
        def __init__(self, owner):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1))
            self.calls, self.owner = [], [owner]

        def forward(self, imgs, targets):
            n = len(imgs)
            self.calls.append(n)
            self.owner[0].base_feat = torch.ones(n, 2048, 19, 32) * self.w
            self.owner[0].ins_feat = torch.ones(n, 40, 256) * self.w
            oh = torch.zeros(n, 40, 9)
            oh[:, :5, 3] = 1.0
            return {"classification": self.w.sum(), "bbox_regression": self.w.sum() * 2, "bbox_ctrness": self.w.sum() * 3,
                    "gt_classes": oh}

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = torch.nn.Module()
            self.backbone.body = torch.nn.Identity()
            self.head = torch.nn.Identity()

    monkeypatch.setattr(dg_fcos, "fcos_resnet50_fpn", lambda **kw: Stub())
    monkeypatch.setattr(dg_fcos.ops, "grad_reverse", lambda x, alpha=0.1: x)
    m = dg_fcos.DGFCOS(9, 2, "dg", [0.5, 0.5, 0.5, 0.05, 0.0001], num_domains=2)
    m.detector = FakeFCOS(m)
    batch = ([torch.zeros(3, 8, 8)] * 2, [torch.zeros(1, 4)] * 2, [torch.ones(1)] * 2, torch.tensor([0, 1]))
    modes, losses = [], []
    for _ in range(16):
        modes.append(m.mode)
        loss = m.training_step(batch)
        assert loss.dim() == 0 and torch.isfinite(loss)
        losses.append(loss)
    assert modes == [0, 1, 0, 2, 0, 3, 0, 4] * 2
    assert m.detector.calls == [2, 2, 2, 1, 1, 2, 1, 1, 2, 1, 1] * 2
    assert float(losses[0]) == 6.0
    # mode 3 as written in DGFCOS.py:210-219
    want = []
    for i in range(2):
        out = m.detector([batch[0][i]], None)
        want.append(F.cross_entropy(m.InsClsPrime[i](m.ins_feat), out["gt_classes"]))
    torch.testing.assert_close(losses[5], 0.05 * torch.mean(torch.stack(want)))
    assert all(not p.requires_grad for p in m.InsCls[0].parameters())        # mode 4 froze them (DGFCOS.py:224-225)


@pytest.fixture
def exact_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.gpu
def test_fcos_training_forward_matches_reference_golden(exact_convs):
    """One training forward of the mirror on the GPU kernels == the reference's fcos.py on CPU: `gt_classes`
    (assignment + `<= 1` rule + area quirk) bit-exact, the three losses to fp32 conv tolerance."""
    gold = np.load(GOLD / "fcos_step.npz")
    model = _mirror().cuda().train()
    imgs, targets = G.fcos_step_inputs()
    out = model([i.cuda() for i in imgs], [{k: v.cuda() for k, v in t.items()} for t in targets])
    shape = tuple(int(v) for v in gold["gt_shape"])
    want = np.unpackbits(gold["gt_classes"])[: int(np.prod(shape))].reshape(shape)
    got = out["gt_classes"].cpu().numpy()
    assert got.shape == shape and np.array_equal(got.astype(np.uint8), want)
    assert not want[1, :, 1:].any()                                           # the 1-GT image: all-zero labels (fcos.py:139)
    for k in ("classification", "bbox_regression", "bbox_ctrness"):
        np.testing.assert_allclose(float(out[k]), float(gold[k]), rtol=2e-4)
    # the fused loss tail and the ATen chain of fcos.py:149-202 agree on the same head outputs
    model.head.fused_loss = False
    out2 = model([i.cuda() for i in imgs], [{k: v.cuda() for k, v in t.items()} for t in targets])
    model.head.fused_loss = True
    for k in ("classification", "bbox_regression", "bbox_ctrness"):
        np.testing.assert_allclose(float(out[k]), float(out2[k]), rtol=1e-5)
    # and the eval path: detections come out of the NMS kernel with TV's layout
    model.eval()
    with torch.no_grad():
        det = model([i.cuda() for i in imgs])
    assert len(det) == len(imgs) and all(d["boxes"].shape[0] == d["scores"].shape[0] <= 100 for d in det)


@pytest.mark.gpu
def test_dgfcos_cycle_on_gpu():
    """BASELINE configs[2] at B=2: one full 8-step mode cycle, finite losses, gradients where DGFCOS.py puts them."""
    from dgod_b200 import dg_fcos, synth
    torch.manual_seed(0)
    B, D = 2, 2
    m = dg_fcos.DGFCOS(9, B, "dg", [0.5, 0.5, 0.5, 0.05, 0.0001], num_domains=D).cuda().train()
    opt = m.configure_optimizer()
    imgs = [i.cuda() for i in synth.random_images(B, 608, 1024, 3)]
    targets, dom = synth.random_targets(B, 6, 608, 1024, 3, n_domains=D)
    batch = (imgs, [t["boxes"].cuda() for t in targets], [t["labels"].cuda() for t in targets], dom.cuda())
    head_w = m.detector.head.classification_head.cls_logits.weight
    for step, mode in enumerate([0, 1, 0, 2, 0, 3, 0, 4]):
        assert m.mode == mode
        loss = m.training_step(batch)
        assert torch.isfinite(loss), (step, mode)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if mode == 0:
            assert head_w.grad is not None and float(head_w.grad.abs().sum()) > 0
        if mode == 1:
            assert m.ImageDA.Conv1.weight.grad is not None and m.InsDA.dc_ip1.weight.grad is not None
            assert m.detector.backbone.fpn.inner_blocks[0][0].weight.grad is not None      # through the GRL
        if mode == 2:
            assert head_w.grad is None and m.InsCls[0].dc_ip1.weight.grad is not None      # detector under no_grad
        if mode == 3:
            assert m.InsClsPrime[0].dc_ip1.weight.grad is not None
            assert m.detector.backbone.fpn.inner_blocks[0][0].weight.grad is not None
        opt.step()
    assert m.mode == 0 and m.sub_mode == 0
    assert m.ins_feat.shape[0] == 1 and m.ins_feat.shape[2] == 256 and m.base_feat.shape[1] == 2048
