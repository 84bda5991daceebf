"""Data-parallel gradient synchronisation for the DG training loops: one process per GPU,
identical replicas, NCCL all-reduce (mean) of the gradients over NVLink / NVSwitch — overlapped with backward.

* Every parameter's `.grad` is a VIEW of one persistent flat buffer, laid out in reverse registration order
  (roughly the order in which backward produces gradients) and cut into buckets.  There is no concatenation
  before the all-reduce and no copy back after it.
* A post-accumulate-grad hook per parameter counts arrivals; when the last gradient a bucket expects for the
  current schedule key (the DG mode) has arrived, the bucket's all-reduce is enqueued asynchronously on NCCL's
  stream while backward continues.  Which parameters receive a gradient depends on the mode (DGFRCNN.py:
  125-199 touches different heads in different modes): the expectation is learned the first time a key is seen
  (everything then goes out in `finish()`), and checked on later steps.
* The layout never depends on the data: a parameter without a gradient on this rank contributes zeros, so ranks
  whose local batches touch different heads (the per-image loops of modes 2-4) still reduce matching buffers.
* The reference's optimizer skips parameters whose `.grad` is None (they get no weight decay).  `finish()`
  reproduces that: a parameter no rank touched in this step has its `.grad` detached to None for the optimizer
  step and re-attached to its view afterwards (`restore()`).

Works on gloo (CPU tests, world size 2) and NCCL.
"""
from __future__ import annotations

from typing import Dict, Hashable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


class GradSync:
    def __init__(self, params: Sequence[Tensor], world_size: int, bucket_bytes: int = 32 << 20,
                 process_group=None):
        self.params: List[Tensor] = [p for p in params if p.requires_grad]
        self.world = int(world_size)
        self.group = process_group
        if not self.params:
            raise ValueError("GradSync: no trainable parameters")
        p0 = self.params[0]
        if any(p.dtype != p0.dtype or p.device != p0.device for p in self.params):
            raise ValueError("GradSync: parameters must share dtype and device")
        order = list(reversed(range(len(self.params))))           # backward produces the last layers' gradients first
        esz = p0.element_size()
        self.offsets: Dict[int, int] = {}
        self.bucket_of: Dict[int, int] = {}
        self.bucket_ranges: List[List[int]] = []                   # [start, end) element ranges of the flat buffer
        off, start, b = 0, 0, 0
        for i in order:
            n = self.params[i].numel()
            n_pad = (n + 31) // 32 * 32                            # keep every view 128-byte aligned
            self.offsets[i] = off
            self.bucket_of[i] = b
            off += n_pad
            if (off - start) * esz >= bucket_bytes:
                self.bucket_ranges.append([start, off])
                start, b = off, b + 1
        if off > start:
            self.bucket_ranges.append([start, off])
        self.flat = torch.zeros(off + len(self.params), dtype=p0.dtype, device=p0.device)
        self.flags = self.flat[off:]                               # one element per parameter: touched on some rank
        self.n_flat = off
        self.views = [self.flat[self.offsets[i]:self.offsets[i] + p.numel()].view_as(p) for i, p in enumerate(self.params)]
        for p, v in zip(self.params, self.views):
            p.grad = v
        self._touched = [False] * len(self.params)
        self._arrived = [0] * len(self.bucket_ranges)
        self._launched: List[Optional[object]] = [None] * len(self.bucket_ranges)
        self._expected: Dict[Hashable, List[int]] = {}
        self._touched_sets: Dict[Hashable, frozenset] = {}
        self._key: Hashable = None
        self._detached: List[int] = []
        self.use_avg = self.world > 1 and dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    # ------------------------------------------------------------------ per-step protocol
    def begin(self, key: Hashable = None) -> None:
        """Call before backward: zeroes the flat buffer (one kernel instead of one per parameter) and arms the
        bucket counters for schedule key `key`."""
        self.restore()
        self.flat.zero_()
        self._key = key
        self._touched = [False] * len(self.params)
        self._arrived = [0] * len(self.bucket_ranges)
        self._launched = [None] * len(self.bucket_ranges)

    def _make_hook(self, i: int):
        def hook(p: Tensor):
            if p.grad is not self.views[i]:                         # autograd replaced the view (first use): fold it back
                self.views[i].copy_(p.grad)
                p.grad = self.views[i]
            if self._touched[i]:
                return
            self._touched[i] = True
            b = self.bucket_of[i]
            self._arrived[b] += 1
            exp = self._expected.get(self._key)
            if self.world > 1 and exp is not None and exp[b] > 0 and self._arrived[b] == exp[b]:
                self._reduce_bucket(b)
        return hook

    def _reduce_bucket(self, b: int) -> None:
        s, e = self.bucket_ranges[b]
        op = dist.ReduceOp.AVG if self.use_avg else dist.ReduceOp.SUM
        self._launched[b] = dist.all_reduce(self.flat[s:e], op=op, group=self.group, async_op=True)

    def finish(self, consistent_across_ranks: bool = True) -> None:
        """Call after backward, before optimizer.step(): sends the buckets that were not sent from the hooks, waits
        for all of them, and hides from the optimizer the parameters that no rank touched.  With
        `consistent_across_ranks=False` (per-image loops whose heads depend on the local domain ids) the touched
        flags are reduced too, at the price of one small device->host read."""
        touched = self._touched
        if self.world > 1:
            late = [b for b in range(len(self.bucket_ranges)) if self._launched[b] is None]
            if not consistent_across_ranks:
                self.flags.copy_(torch.tensor([1.0 if t else 0.0 for t in touched], dtype=self.flat.dtype), non_blocking=False)
            for b in late:
                self._reduce_bucket(b)
            work_flags = None
            if not consistent_across_ranks:
                work_flags = dist.all_reduce(self.flags, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            for w in self._launched:
                if w is not None:
                    w.wait()
            if not self.use_avg:
                self.flat[:self.n_flat].div_(self.world)
            if work_flags is not None:
                work_flags.wait()
                touched = [v > 0 for v in self.flags.tolist()]
            key = self._key
            counts = [0] * len(self.bucket_ranges)
            for i, t in enumerate(self._touched):
                if t:
                    counts[self.bucket_of[i]] += 1
            known = self._expected.get(key)
            if known is None or known != counts:
                # first time this key is seen (or the touched set changed): remember it; buckets whose expectation
                # was wrong went out late, which costs overlap but never correctness
                self._expected[key] = counts
        self._detached = [i for i, t in enumerate(touched) if not t]
        for i in self._detached:
            self.params[i].grad = None

    def restore(self) -> None:
        """Re-attach the gradient views that `finish()` hid from the optimizer."""
        for i in self._detached:
            self.params[i].grad = self.views[i]
        self._detached = []


def allreduce_gradients(params, world_size: int) -> None:
    """Synchronous fallback: ONE flat all-reduce (mean) over every trainable parameter, zeros where this rank has
    no gradient, so that the layout never depends on the local data (dgod_b200.dg.allreduce_gradients)."""
    if world_size <= 1:
        return
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world_size)
    off = 0
    for p in ps:
        n = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
