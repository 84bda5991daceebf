"""Loader for the UNMODIFIED reference modules installed in baseline/_ref (oracle/install_ref.py).
TEST / BENCH INFRASTRUCTURE ONLY.

The reference imports pytorch_lightning and torchmetrics (DGcommon.py:9-10), which are not installed, asks for
COCO-pretrained weights (DGFRCNN.py:81, DGFCOS.py:119 — no network) and hard-codes `.cuda()` / `.to(device=0)`
in its training_step (DGFRCNN.py:112-117,150,154; DGFCOS.py:156-162).  This module supplies, WITHOUT editing a
line of the reference:
  * stub modules for the two missing imports (SURVEY.md §8c recipe);
  * a wrapper of the two detector factories that forces random init but keeps what `pretrained=True` implies for
    the training graph: trainable_backbone_layers=3, i.e. conv1 / layer1 frozen (fasterrcnn.py:308-317);
  * on a machine without CUDA, `.cuda()` and `.to(device=0)` become no-ops so that the very same training_step
    runs on the host cores (the CPU baseline of bench.py).
"""
from __future__ import annotations

import sys
import types
import warnings
from pathlib import Path

import torch
from torch import nn

ROOT = Path(__file__).resolve().parent.parent
REF_DIRS = [ROOT / "baseline" / "_ref", Path("/root/reference")]


def available() -> bool:
    return any((d / "DGFRCNN.py").exists() for d in REF_DIRS)


def _stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        core = types.ModuleType("pytorch_lightning.core")
        mod = types.ModuleType("pytorch_lightning.core.module")

        class LightningModule(nn.Module):
            def log(self, *a, **k):
                pass

        mod.LightningModule = LightningModule
        core.module = mod
        pl.core = core
        pl.LightningModule = LightningModule
        sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.core": core, "pytorch_lightning.core.module": mod})
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        det = types.ModuleType("torchmetrics.detection")

        class MeanAveragePrecision:
            def __init__(self, *a, **k):
                pass

            def update(self, *a, **k):
                pass

            def compute(self):
                return {"map_50": torch.tensor(0.0), "map_per_class": torch.tensor(0.0)}

            def reset(self):
                pass

        det.MeanAveragePrecision = MeanAveragePrecision
        tm.detection = det
        sys.modules.update({"torchmetrics": tm, "torchmetrics.detection": det})


def _freeze_like_trainable_3(backbone_body: nn.Module):
    """What resnet_fpn_backbone(trainable_layers=3) does (TV backbone_utils.py): only layer2-4 train."""
    for name, p in backbone_body.named_parameters():
        if not any(name.startswith(l) for l in ("layer4", "layer3", "layer2")):
            p.requires_grad_(False)


def load(cpu_shims: bool = None):
    """Imports the reference; returns the modules namespace (fasterrcnn, fcos, DGcommon, DGFRCNN, DGFCOS)."""
    ref = next((d for d in REF_DIRS if (d / "DGFRCNN.py").exists()), None)
    if ref is None:
        raise RuntimeError("the reference is not installed: run `python -m oracle.install_ref` where /root/reference exists")
    _stubs()
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    warnings.filterwarnings("ignore")
    import fasterrcnn
    import fcos
    if not getattr(fasterrcnn, "_dgod_offline", False):
        f_orig, c_orig = fasterrcnn.fasterrcnn_resnet50_fpn, fcos.fcos_resnet50_fpn

        def frcnn_factory(pretrained=False, progress=True, num_classes=91, pretrained_backbone=True,
                          trainable_backbone_layers=3, **kw):
            m = f_orig(pretrained=False, progress=progress, num_classes=num_classes, pretrained_backbone=False,
                       trainable_backbone_layers=trainable_backbone_layers, **kw)
            if pretrained or pretrained_backbone:        # the graph the reference trains: frozen stem (fasterrcnn.py:311-317)
                if trainable_backbone_layers == 3:
                    _freeze_like_trainable_3(m.backbone.body)
            return m

        def fcos_factory(*a, **kw):
            kw["pretrained_backbone"] = False
            return c_orig(*a, **kw)

        fasterrcnn.fasterrcnn_resnet50_fpn = frcnn_factory
        fcos.fcos_resnet50_fpn = fcos_factory
        fasterrcnn._dgod_offline = True
    import DGcommon
    import DGFRCNN
    import DGFCOS
    if cpu_shims is None:
        cpu_shims = not torch.cuda.is_available()
    if cpu_shims and not getattr(torch.Tensor, "_dgod_cpu_shims", False):
        to_orig = torch.Tensor.to

        def to(self, *a, **k):
            if k.get("device", None) == 0:
                k = {kk: v for kk, v in k.items() if kk != "device"}
                if not a and not k:
                    return self
            return to_orig(self, *a, **k)

        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.Tensor.to = to
        torch.Tensor._dgod_cpu_shims = True
    return types.SimpleNamespace(fasterrcnn=fasterrcnn, fcos=fcos, DGcommon=DGcommon, DGFRCNN=DGFRCNN, DGFCOS=DGFCOS)


def build_dgfrcnn(n_classes: int, batch_size: int, exp: str, reg_weights, num_domains: int):
    """The reference's DGFRCNN (DGFRCNN.py:73-201) exactly as train_driving_dg.py constructs it."""
    R = load()
    return R.DGFRCNN.DGFRCNN(n_classes, batch_size, exp, list(reg_weights), None, [None] * num_domains)


def build_dgfcos(n_classes: int, batch_size: int, exp: str, reg_weights, num_domains: int):
    R = load()
    return R.DGFCOS.DGFCOS(n_classes, batch_size, exp, list(reg_weights), None, [None] * num_domains)
