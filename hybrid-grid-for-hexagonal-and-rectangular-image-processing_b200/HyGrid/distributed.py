"""Multi-GPU plumbing of the path: one process per GPU, ``torch.distributed``.

* Resampling, layout conversion, pooling, conv forward / data gradient are independent per image: the batch
  is cut into contiguous shards (``shard_range`` / ``shard_batch``) and **no collective** is involved.
* The only exchange step is the hex-conv weight / bias gradient (plus norm affine gradients in a training
  step): every layer's gradient lives in ONE flat fp32 bucket (``FlatGradBucket``), so a training step issues a
  single all-reduce (NCCL over NVLink / NVSwitch on the GPU box; gloo in the CPU tests) instead of one
  latency-bound collective per tiny tensor (64*64*7+64 floats = 115 KB per layer).

The reference has no distributed code at all; this module is new surface.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

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


class FlatGradBucket:
    """All gradients of ``params`` as views into one flat buffer, reduced with a single collective.

    ``bucket = FlatGradBucket(model.parameters())`` re-points every ``p.grad`` at a slice of ``bucket.flat``
    (autograd then accumulates in place, so the hex-conv weight-gradient partials land in the bucket without
    a gather copy).  After ``loss.backward()``: ``bucket.all_reduce()`` (sum, then divide by the world size when
    ``average``); ``bucket.zero_()`` replaces ``optimizer.zero_grad()``."""

    def __init__(self, params: Iterable[torch.nn.Parameter], dtype: torch.dtype = torch.float32):
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
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        self.attach()

    def view(self, i: int) -> torch.Tensor:
        p = self.params[i]
        return self.flat[self.offsets[i]:self.offsets[i] + p.numel()].view_as(p)

    def attach(self) -> None:
        for i, p in enumerate(self.params):
            v = self.view(i)
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v if v.dtype == p.dtype else None
        self._mixed = any(p.dtype != self.flat.dtype for p in self.params)

    def zero_(self) -> None:
        self.flat.zero_()

    def gather(self) -> None:
        """Copy gradients that autograd re-allocated (or that have another dtype) back into the bucket."""
        for i, p in enumerate(self.params):
            v = self.view(i)
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                if v.dtype == p.dtype:
                    p.grad = v

    def all_reduce(self, group=None, average: bool = True, async_op: bool = False):
        self.gather()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        if async_op:
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
            return _Scaled(work, self.flat, 1.0 / world if average else 1.0)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(world)
        return None

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()


class _Scaled:
    def __init__(self, work, flat, scale):
        self.work, self.flat, self.scale = work, flat, scale

    def wait(self):
        self.work.wait()
        if self.scale != 1.0:
            self.flat.mul_(self.scale)
