"""Multi-GPU plumbing of the path: one process per GPU, ``torch.distributed``.

* Resampling, layout conversion, pooling, conv forward / data gradient are independent per image: the batch
  is cut into contiguous shards (``shard_range`` / ``shard_batch``) and **no collective** is involved.
* The only exchange step is the hex-conv weight / bias gradient (plus norm affine gradients in a training
  step): every layer's gradient lives in ONE flat fp32 bucket (``FlatGradBucket``).  The bucket is cut into a few
  contiguous *groups* of consecutive parameters; as soon as the backward pass has produced the last gradient of a
  group, that slice is all-reduced asynchronously (NCCL over NVLink / NVSwitch on the GPU box -- the collective runs
  on NCCL's own stream, which waits for the weight-gradient kernel that just landed, while the backward of the earlier
  layers keeps running on the compute stream; gloo in the CPU tests).  ``finish()`` joins before the optimizer step.
  Messages are tiny (64*64*7+64 floats = 115 KB per layer) and latency-bound, so a few groups, not one collective per
  tensor.
* The hex-conv weight-gradient kernel accumulates into its destination (include/hygrid_b200.h, hg_hexconv_wgrad), so
  the bucket hands every conv parameter its slice as a *gradient sink* (``Parameter._hg_grad_sink``): the kernel's
  partial sums land in the all-reduce buffer directly -- no zero-filled temporary, no autograd ``add_``.

The reference has no distributed code at all; this module is new surface.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_batch", "FlatGradBucket"]


def _rank_world(rank, world, group=None) -> Tuple[int, int]:
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(group), dist.get_world_size(group)
        return 0, 1
    return int(rank), int(world)


def shard_range(n: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous ``[start, stop)`` of ``n`` units owned by ``rank``: sizes differ by at most one, earlier
    ranks take the remainder, the shards tile ``range(n)`` exactly."""
    rank, world = _rank_world(rank, world)
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(int(n), world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None, dim: int = 0) -> torch.Tensor:
    """This rank's contiguous slice of the batch dimension (a view, no copy)."""
    a, b = shard_range(x.shape[dim], rank, world)
    return x.narrow(dim, a, b - a)


class GradSink:
    """What a conv parameter carries as ``_hg_grad_sink``: the fp32 slice of the bucket its gradient accumulates into and
    the callback that tells the bucket the slice is complete for this backward pass."""
    __slots__ = ("view", "_bucket", "_index")

    def __init__(self, view, bucket, index):
        self.view, self._bucket, self._index = view, bucket, index

    def landed(self):
        self._bucket._ready(self._index)


class FlatGradBucket:
    """All gradients of ``params`` as views into one flat buffer, reduced group by group while backward still runs.

    ``bucket = FlatGradBucket(model.parameters(), groups=3)`` re-points every ``p.grad`` at a slice of ``bucket.flat``
    (autograd then accumulates in place; hex-conv kernels write there directly through the gradient sink).  Per step::

        bucket.zero_()                 # replaces optimizer.zero_grad()
        loss.backward()                # group g is all-reduced as soon as its last gradient has landed (overlap=True)
        bucket.finish()                # join the collectives (launches whatever has not been launched), average
        optimizer.step()

    ``bucket.all_reduce()`` keeps the round-1 behaviour (one blocking collective over the whole buffer after backward)
    for callers that built the bucket with ``overlap=False``.  Parameters whose dtype differs from the bucket's (e.g.
    bfloat16 parameters in a float32 bucket) keep their own ``.grad`` tensor; it is copied into the bucket when it lands
    and the reduced value is written back by ``finish()`` / ``all_reduce()``."""

    def __init__(self, params: Iterable[torch.nn.Parameter], dtype: torch.dtype = torch.float32, groups=1,
                 overlap: bool = False, group=None, average: bool = True):
        """``groups``: an int (that many runs of consecutive parameters, cut by tensor count so that a run ends with a
        layer's last tensor rather than in the middle of the largest one) or a sequence of run lengths, e.g.
        ``[3, 3, 5]`` = first three tensors, next three, last five.  Backward produces gradients last layer first, so
        the LAST run is reduced first, while the earlier layers' backward is still running."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        if any(p.device != dev for p in self.params):
            raise ValueError("all parameters of a bucket must live on one device")
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 31) // 32 * 32          # 128-byte aligned slices
        self.total = total
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        self.group, self.average, self.overlap = group, average, bool(overlap)
        if isinstance(groups, int):
            self.bounds = self._cut(max(1, min(int(groups), len(self.params))))
        else:
            sizes = [int(v) for v in groups]
            if sum(sizes) != len(self.params) or any(v < 1 for v in sizes):
                raise ValueError(f"group sizes {sizes} do not tile the {len(self.params)} trainable parameters")
            ends = [sum(sizes[:i + 1]) for i in range(len(sizes))]
            self.bounds = list(zip([0] + ends[:-1], ends))
        self.group_of = [g for g, (a, b) in enumerate(self.bounds) for _ in range(a, b)]
        self.trace: List[Tuple[str, int]] = []           # ("ready", param index) / ("launch", group index), per step
        self._hooks = []
        self._views: List[Optional[torch.Tensor]] = [None] * len(self.params)
        self._reset_step()
        self.attach()

    # ------------------------------------------------------------------ layout
    def _cut(self, n: int) -> List[Tuple[int, int]]:
        """``n`` runs of consecutive parameters with (almost) equal numbers of TENSORS: [(first, last+1), ...].  The
        messages are latency-bound (a few hundred KB in total), so balancing bytes buys nothing; what matters is that a
        run is complete -- and its collective launched -- as early in the backward pass as possible."""
        count = len(self.params)
        base, rem = divmod(count, n)
        cuts, start = [], 0
        for g in range(n):
            size = base + (1 if g >= n - rem else 0)          # the later (first-ready) runs take the remainder
            if size:
                cuts.append((start, start + size))
                start += size
        return cuts

    def view(self, i: int) -> torch.Tensor:
        v = self._views[i] if i < len(self._views) else None      # cached: _ready() runs once per parameter and step
        if v is None:
            p = self.params[i]
            v = self.flat[self.offsets[i]:self.offsets[i] + p.numel()].view_as(p)
            if i < len(self._views):
                self._views[i] = v
        return v

    def _slice(self, g: int) -> torch.Tensor:
        a, b = self.bounds[g]
        end = self.offsets[b] if b < len(self.params) else self.total
        return self.flat[self.offsets[a]:end]

    def _native(self, i: int) -> bool:
        return self.params[i].dtype == self.flat.dtype

    def attach(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for i, p in enumerate(self.params):
            v = self.view(i)
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            if self._native(i):
                p.grad = v
                p._hg_grad_sink = GradSink(v, self, i) if self.flat.dtype == torch.float32 else None
            else:
                p._hg_grad_sink = None
            self._hooks.append(p.register_post_accumulate_grad_hook(lambda _p, i=i: self._ready(i)))
        self._mixed = [i for i in range(len(self.params)) if not self._native(i)]

    def detach(self) -> None:
        """Remove hooks and gradient sinks (the parameters keep their current ``.grad`` views)."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for p in self.params:
            p._hg_grad_sink = None

    # ------------------------------------------------------------------ step
    def _reset_step(self) -> None:
        self._pending = [b - a for a, b in self.bounds]
        self._seen = [False] * len(self.params)
        self._works = [None] * len(self.bounds)
        self.trace = []

    def zero_(self) -> None:
        self.flat.zero_()
        for i in self._mixed:
            self.params[i].grad = None
        self._reset_step()

    def _ready(self, i: int) -> None:
        """The gradient of parameter ``i`` is complete for this backward pass (autograd hook or gradient sink)."""
        if self._seen[i]:
            return
        self._seen[i] = True
        self.trace.append(("ready", i))
        p = self.params[i]
        v = self.view(i)
        if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
            v.copy_(p.grad)                                # another dtype, or autograd re-allocated it
            if self._native(i):
                p.grad = v
        g = self.group_of[i]
        self._pending[g] -= 1
        if self.overlap and self._pending[g] == 0:
            self._launch(g)

    def _distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _launch(self, g: int) -> None:
        if self._works[g] is not None:
            return
        self.trace.append(("launch", g))
        if not self._distributed():
            self._works[g] = "local"
            return
        t = self._slice(g)
        if self.average and dist.get_backend(self.group) == "nccl":
            self._works[g] = (dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True), 1.0)
        else:
            w = dist.get_world_size(self.group)
            self._works[g] = (dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True),
                              1.0 / w if self.average else 1.0)

    def gather(self) -> None:
        """Copy gradients that autograd re-allocated (or that have another dtype) into the bucket."""
        for i, p in enumerate(self.params):
            v = self.view(i)
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr() and not self._seen[i]:
                v.copy_(p.grad)
                if self._native(i):
                    p.grad = v

    def _scatter_mixed(self) -> None:
        for i in self._mixed:
            p = self.params[i]
            if p.grad is None:
                p.grad = self.view(i).to(p.dtype)
            else:
                p.grad.copy_(self.view(i))

    def finish(self) -> None:
        """Join the step's collectives.  Groups whose gradients never all arrived (unused parameters) are reduced now;
        the groups must be launched in the same order on every rank, which holds because every rank runs the same
        backward graph."""
        self.gather()
        for g in reversed(range(len(self.bounds))):
            self._launch(g)
        scale = []
        for g, w in enumerate(self._works):
            if isinstance(w, tuple):
                w[0].wait()
                if w[1] != 1.0:
                    scale.append(g)
        if len(scale) == len(self.bounds):
            self.flat.mul_(self._works[0][1])
        else:
            for g in scale:
                self._slice(g).mul_(self._works[g][1])
        self._scatter_mixed()
        self._pending = [b - a for a, b in self.bounds]
        self._seen = [False] * len(self.params)
        self._works = [None] * len(self.bounds)

    def all_reduce(self, group=None, average: bool = True, async_op: bool = False):
        """One collective over the whole buffer after backward (no overlap)."""
        self.gather()
        self._seen = [False] * len(self.params)
        self._pending = [b - a for a, b in self.bounds]
        group = group if group is not None else self.group
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            self._scatter_mixed()
            return None
        world = dist.get_world_size(group)
        if async_op:
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
            return _Scaled(work, self, 1.0 / world if average else 1.0)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(world)
        self._scatter_mixed()
        return None

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def group_bytes(self) -> Sequence[int]:
        return [self._slice(g).numel() * self.flat.element_size() for g in range(len(self.bounds))]


class _Scaled:
    def __init__(self, work, bucket, scale):
        self.work, self.bucket, self.scale = work, bucket, scale

    def wait(self):
        self.work.wait()
        if self.scale != 1.0:
            self.bucket.flat.mul_(self.scale)
        self.bucket._scatter_mixed()
