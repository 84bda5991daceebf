"""Host-side utilities shared by bench.py's two arms and the tests (no kernels here)."""
from __future__ import annotations

import torch
from torch import nn


@torch.no_grad()
def calibrate_frozen_bn(backbone: nn.Module, images: torch.Tensor) -> int:
    """Data-dependent initialisation of the FrozenBatchNorm2d statistics of a randomly initialised
    backbone: one forward pass in which every frozen BN layer first sets running_mean / running_var
    to the statistics of its own input.

    The reference trains from COCO-pretrained weights (DGFRCNN.py:81), which cannot be downloaded
    here.  A random-init ResNet-50 whose frozen BN layers are the identity (mean 0, var 1) has
    activations that grow by orders of magnitude with depth and diverges to NaN on the first
    SGD step at the reference's learning rate, after which the RPN proposes nothing and the timed
    work collapses.  Calibrated statistics give the activations the unit scale they have in a trained
    network without changing the architecture or the amount of work.  Returns the number of layers."""
    from torchvision.ops.misc import FrozenBatchNorm2d
    hooks, count = [], 0

    def pre(mod, inputs):
        x = inputs[0]
        mod.running_mean.copy_(x.mean(dim=(0, 2, 3)))
        mod.running_var.copy_(x.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-6))

    for m in backbone.modules():
        if isinstance(m, FrozenBatchNorm2d):
            hooks.append(m.register_forward_pre_hook(pre))
            count += 1
    was_training = backbone.training
    backbone.eval()
    backbone(images)
    backbone.train(was_training)
    for h in hooks:
        h.remove()
    return count
