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


# ------------------------------------------------------------------ frozen-BN folding (host side, PyTorch/cuDNN only)
class _FoldedConv2d(nn.Conv2d):
    """conv followed by a FrozenBatchNorm2d, evaluated as ONE cuDNN convolution:
    bn(conv(x, w)) = conv(x, w * s) + (b - m * s) with s = gamma / sqrt(var + eps), per output channel.
    The parameter stays `weight` (same state_dict key, same weight decay, gradients flow through the
    product), the BN buffers stay in the BN module; only the two elementwise passes over the
    activation (x*s, +shift) and their backward disappear."""

    def forward(self, x):
        bn = self._folded_bn[0]
        scale = bn.weight * (bn.running_var + bn.eps).rsqrt()
        shift = bn.bias - bn.running_mean * scale
        return self._conv_forward(x, self.weight * scale.view(-1, 1, 1, 1), shift)


def fold_frozen_bn(module: nn.Module) -> int:
    """Rebinds every (Conv2d without bias, FrozenBatchNorm2d) sibling pair that torchvision's ResNet applies
    back to back (stem, Bottleneck conv1-3, downsample) so that the BN becomes the conv's epilogue.
    Call after `calibrate_frozen_bn`.  State-dict keys and parameters are unchanged.  Returns the
    number of folded pairs; `unfold_frozen_bn` restores the stock modules."""
    from torchvision.ops.misc import FrozenBatchNorm2d

    class _Identity(FrozenBatchNorm2d):
        def forward(self, x):
            return x

    n = 0
    for parent in module.modules():
        prev = None
        for child in parent.children():
            if (type(child) is FrozenBatchNorm2d and type(prev) is nn.Conv2d and prev.bias is None
                    and prev.out_channels == child.weight.numel()):
                prev.__class__ = _FoldedConv2d
                prev._folded_bn = [child]          # a list: not registered as a sub-module
                child.__class__ = _Identity
                n += 1
            prev = child
    return n


def unfold_frozen_bn(module: nn.Module) -> None:
    from torchvision.ops.misc import FrozenBatchNorm2d
    for m in module.modules():
        if isinstance(m, _FoldedConv2d):
            m._folded_bn[0].__class__ = FrozenBatchNorm2d
            del m._folded_bn
            m.__class__ = nn.Conv2d
