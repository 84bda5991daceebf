"""Seeded synthetic inputs shared by the tests, the oracle runs and bench.py.

Everything is generated on the CPU with an explicit `torch.Generator`, so the CPU oracle and the
GPU path see bit-identical inputs.  Image / target layout follows what the reference's dataset
yields (DrivingDataset.py:71): image float [3,H,W] in [0,1), boxes [n,4] xyxy, labels [n] in
1..8, domain id.
"""
from __future__ import annotations

import torch


def gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def random_boxes(n: int, height: float, width: float, g: torch.Generator,
                 log_size=(2.5, 4.0), min_side: float = 1.0) -> torch.Tensor:
    """Centre uniform over the image, sqrt(area) log-uniform in exp(2.5)..exp(6.5) px (12..665),
    aspect exp((U-.5)*1.4); clipped to the image with sides of at least `min_side`."""
    cx = torch.rand(n, generator=g) * width
    cy = torch.rand(n, generator=g) * height
    size = torch.exp(torch.rand(n, generator=g) * log_size[1] + log_size[0])
    aspect = torch.exp((torch.rand(n, generator=g) - 0.5) * 1.4)
    w, h = size * torch.sqrt(aspect), size / torch.sqrt(aspect)
    x1 = (cx - w / 2).clamp(0, width - min_side)
    y1 = (cy - h / 2).clamp(0, height - min_side)
    x2 = torch.maximum((cx + w / 2).clamp(0, width), x1 + min_side)
    y2 = torch.maximum((cy + h / 2).clamp(0, height), y1 + min_side)
    return torch.stack([x1, y1, x2, y2], dim=1).float()


def random_targets(batch: int, n_gt: int, height: int, width: int, seed: int, n_domains: int = 2):
    """List of per-image dicts {boxes, labels} plus the domain id vector (i mod D)."""
    targets = []
    for i in range(batch):
        g = gen(seed * 1000 + i)
        boxes = random_boxes(n_gt, height, width, g)
        labels = torch.randint(1, 9, (n_gt,), generator=g, dtype=torch.int64)
        targets.append({"boxes": boxes, "labels": labels})
    domains = torch.arange(batch, dtype=torch.int64) % n_domains
    return targets, domains


def random_images(batch: int, height: int, width: int, seed: int):
    g = gen(seed)
    return [torch.rand(3, height, width, generator=g) for _ in range(batch)]


def fpn_shapes(height: int, width: int, strides=(4, 8, 16, 32)):
    """Feature grid of a padded image at each stride (ceil division, as the backbone produces)."""
    return [((height + s - 1) // s, (width + s - 1) // s) for s in strides]


def random_features(batch: int, channels: int, height: int, width: int, seed: int,
                    strides=(4, 8, 16, 32), dtype=torch.float32):
    g = gen(seed)
    return [torch.randn(batch, channels, h, w, generator=g).to(dtype)
            for (h, w) in fpn_shapes(height, width, strides)]


def rois_from_boxes(boxes_per_image) -> torch.Tensor:
    """[K,5] (batch index, x1, y1, x2, y2) — TV ops/poolers.py:87-95."""
    out = []
    for i, b in enumerate(boxes_per_image):
        out.append(torch.cat([torch.full((len(b), 1), float(i), dtype=b.dtype), b], dim=1))
    return torch.cat(out, dim=0) if out else torch.zeros(0, 5)


def distinct_scores(n: int, g: torch.Generator) -> torch.Tensor:
    """Scores in (0,1) that are pairwise distinct in fp32 (torch.rand has only 2^24 values, so
    ties appear from a few thousand draws on; tie order is implementation-defined upstream)."""
    perm = torch.randperm(n, generator=g).double()
    jitter = torch.rand(n, generator=g, dtype=torch.float64) * 0.5
    s = ((perm + 0.25 + jitter) / max(n, 1)).float()
    assert len(torch.unique(s)) == n
    return s
