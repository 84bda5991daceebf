"""Data-parallel gradient synchronisation for the DG training loops: one process per GPU,
identical replicas, NCCL all-reduce (mean) of the gradients over NVLink / NVSwitch — overlapped with backward.

* One persistent flat buffer holds every trainable parameter's gradient, laid out in reverse registration order
  (roughly the order in which backward produces gradients) and cut into buckets; each parameter has a VIEW into it
  with the parameter's own strides (channels_last weights get channels_last views).
* Backward runs with `.grad = None`, so autograd hands every parameter its gradient tensor without an extra
  accumulate kernel.  A post-accumulate-grad hook per parameter counts arrivals; when the last gradient a bucket
  expects for the current schedule key (the DG mode) has arrived, the bucket is packed with ONE multi-tensor copy and
  its all-reduce is enqueued asynchronously on NCCL's stream while backward continues.  Which parameters receive a
  gradient depends on the mode (DGFRCNN.py:125-199 touches different heads in different modes): the expectation is
  learned the first time a key is seen (everything then goes out in `finish()`).
* The layout never depends on the data: a parameter without a gradient on this rank contributes zeros, so ranks whose
  local batches touch different heads (the per-image loops of modes 2-4) still reduce matching buffers.
* After `finish()` a touched parameter's `.grad` IS its view of the reduced buffer (no copy back); a parameter no
  rank touched keeps `.grad = None`, so the optimizer skips it exactly as it does for the reference (no weight decay
  on heads the step did not use).

Works on gloo (CPU tests, world size 2) and NCCL.
"""
from __future__ import annotations

from typing import Dict, Hashable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def _view_like(flat: Tensor, p: Tensor) -> Tensor:
    """A view of the 1-D `flat` (p.numel() elements) with p's shape AND p's strides when p is dense in some permutation
    of its dimensions (contiguous, channels_last, ...); contiguous otherwise."""
    if p.is_contiguous() or p.dim() < 2:
        return flat.view(p.shape)
    order = sorted(range(p.dim()), key=lambda d: (-p.stride(d), d))           # dimensions from outermost to innermost
    shape = [p.shape[d] for d in order]
    v = flat.view(shape)
    inv = [order.index(d) for d in range(p.dim())]
    v = v.permute(inv)
    return v if v.stride() == p.stride() else flat.view(p.shape)


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
        offsets: Dict[int, int] = {}
        self.bucket_of: Dict[int, int] = {}
        self.bucket_ranges: List[List[int]] = []                   # [start, end) element ranges of the flat buffer
        self.members: List[List[int]] = [[]]
        off, start = 0, 0
        for i in order:
            n = self.params[i].numel()
            offsets[i] = off
            self.bucket_of[i] = len(self.bucket_ranges)
            self.members[-1].append(i)
            off += (n + 31) // 32 * 32                             # keep every view 128-byte aligned
            if (off - start) * esz >= bucket_bytes:
                self.bucket_ranges.append([start, off])
                self.members.append([])
                start = off
        if off > start:
            self.bucket_ranges.append([start, off])
        else:
            self.members.pop()
        self.flat = torch.zeros(off + len(self.params), dtype=p0.dtype, device=p0.device)
        self.flags = self.flat[off:]                               # one element per parameter: touched on some rank
        self.n_flat = off
        self.views = [_view_like(self.flat[offsets[i]:offsets[i] + p.numel()], p) for i, p in enumerate(self.params)]
        self._touched = [False] * len(self.params)
        self._arrived = [0] * len(self.bucket_ranges)
        self._launched: List[Optional[object]] = [None] * len(self.bucket_ranges)
        self._packed = [False] * len(self.bucket_ranges)
        self._expected: Dict[Hashable, List[int]] = {}
        self._key: Hashable = None
        self.use_avg = self.world > 1 and dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        if self.use_avg:
            # all-reduce kernels now run during backward: the RoIAlign backward claims its work items instead of dealing them
            # (a persistent grid with a static deal waits for the CTAs whose SMs NCCL still holds: 342 vs 281 us at N=2)
            import os
            from . import ops
            if "DGOD_BWD_ALGO" not in os.environ and ops.BACKWARD_ALGO == 0:
                ops.BACKWARD_ALGO = 5
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    # ------------------------------------------------------------------ per-step protocol
    def begin(self, key: Hashable = None) -> None:
        """Call before backward: clears the flat buffer (one kernel) and every `.grad`, arms the bucket counters for
        schedule key `key`."""
        self.flat.zero_()
        for p in self.params:
            p.grad = None
        self._key = key
        self._touched = [False] * len(self.params)
        self._arrived = [0] * len(self.bucket_ranges)
        self._launched = [None] * len(self.bucket_ranges)
        self._packed = [False] * len(self.bucket_ranges)

    def _make_hook(self, i: int):
        b = self.bucket_of[i]

        def hook(p: Tensor):
            if self._touched[i]:
                return
            self._touched[i] = True
            self._arrived[b] += 1
            exp = self._expected.get(self._key)
            if self.world > 1 and exp is not None and exp[b] > 0 and self._arrived[b] == exp[b]:
                self._send_bucket(b)
        return hook

    def _send_bucket(self, b: int) -> None:
        """Packs the gradients that arrived for bucket b into its slice of the flat buffer (one multi-tensor copy) and
        starts its all-reduce."""
        src = [self.params[i].grad for i in self.members[b] if self._touched[i] and self.params[i].grad is not None]
        dst = [self.views[i] for i in self.members[b] if self._touched[i] and self.params[i].grad is not None]
        if src:
            torch._foreach_copy_(dst, src)
        self._packed[b] = True
        if self.world > 1:
            s, e = self.bucket_ranges[b]
            op = dist.ReduceOp.AVG if self.use_avg else dist.ReduceOp.SUM
            self._launched[b] = dist.all_reduce(self.flat[s:e], op=op, group=self.group, async_op=True)

    def finish(self, consistent_across_ranks: bool = True) -> None:
        """Call after backward, before optimizer.step(): sends the buckets that were not sent from the hooks, waits for
        all of them and points every touched parameter's `.grad` at its view of the reduced buffer.  With
        `consistent_across_ranks=False` (per-image loops whose heads depend on the local domain ids) the touched flags
        are reduced too, at the price of one small device->host read."""
        touched = list(self._touched)
        work_flags = None
        if self.world > 1 and not consistent_across_ranks:
            self.flags.copy_(torch.tensor([1.0 if t else 0.0 for t in touched], dtype=self.flat.dtype))
        for b in range(len(self.bucket_ranges)):
            if not self._packed[b]:
                self._send_bucket(b)
        if self.world > 1:
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
            counts = [0] * len(self.bucket_ranges)
            for i, t in enumerate(self._touched):
                if t:
                    counts[self.bucket_of[i]] += 1
            if self._expected.get(self._key) != counts:
                # first time this key is seen (or the touched set changed): remember it; buckets whose expectation was
                # wrong went out late, which costs overlap but never correctness
                self._expected[self._key] = counts
        for i, t in enumerate(touched):
            self.params[i].grad = self.views[i] if t else None

    def restore(self) -> None:
        """Kept for API compatibility: nothing to undo (untouched parameters simply have `.grad = None`)."""


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
