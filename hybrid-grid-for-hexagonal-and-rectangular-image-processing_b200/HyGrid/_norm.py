"""Batch normalisation (+ fused ReLU) of ``HexConvModule`` on the library's streaming kernels.

The reference's ``HexConvModule`` (HexModules.py:146-288) is conv -> norm -> act with the norm layer built by mmcv
(``torch.nn.BatchNorm2d`` for ``dict(type='BN')``, HexModules.py:57-76).  The module object, its parameters, buffers
and ``state_dict`` keys stay exactly torch's; only the arithmetic of the float32 CUDA forward / backward is replaced
(``hg_bn_stats`` / ``hg_bn_apply`` / ``hg_bn_bwd_reduce`` / ``hg_bn_bwd_apply``), optionally with the following ReLU
fused into the same pass.  Semantics follow ``F.batch_norm``: batch statistics (biased variance) in training mode
or when no running statistics are tracked, running statistics otherwise; running averages updated with
``momentum`` (``None`` = cumulative average) and the unbiased variance.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import _native as nv


def bn_supported(bn: nn.Module, x: torch.Tensor) -> bool:
    """True when ``bn(x)`` can run on the library kernels (else the caller keeps torch's own BatchNorm2d)."""
    return (type(bn) is nn.BatchNorm2d and isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 4
            and x.dtype == torch.float32 and x.numel() > 0
            and (bn.weight is None or bn.weight.dtype == torch.float32))


class _BatchNormReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, batches, use_batch_stats, momentum, eps, relu):
        """``momentum``: None = do not touch the running statistics, < 0 = cumulative average over ``batches``
        (torch's ``momentum=None``), else the blend factor.  One library call per mode (``hg_bn_train_fwd`` zeroes its
        scratch, bumps ``num_batches_tracked`` and updates the running statistics inside the kernels)."""
        x = x.contiguous()
        N, Cc, H, W = x.shape
        HW = H * W
        dev, st = x.device, nv.stream_ptr(x.device)
        y = torch.empty_like(x)
        mean = torch.empty(Cc, dtype=torch.float32, device=dev)
        rstd = torch.empty(Cc, dtype=torch.float32, device=dev)
        w = weight.detach().contiguous() if weight is not None else None
        b = bias.detach().contiguous() if bias is not None else None
        if use_batch_stats:
            if N * HW <= 1:
                raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(x.shape)}")
            sums = torch.empty(2 * Cc, dtype=torch.float64, device=dev)
            update = momentum is not None and running_mean is not None
            fused = update and running_mean.dtype == torch.float32 and running_var.dtype == torch.float32 \
                and running_mean.is_contiguous() and running_var.is_contiguous() \
                and (batches is None or batches.dtype == torch.int64)
            nv.call("hg_bn_train_fwd", nv.ptr(x), nv.ptr(y), nv.ptr(sums), nv.ptr(w), nv.ptr(b), nv.ptr(mean), nv.ptr(rstd),
                    nv.ptr(running_mean) if fused else None, nv.ptr(running_var) if fused else None,
                    nv.ptr(batches) if fused and batches is not None else None, float(momentum) if fused else 0.0,
                    N, Cc, HW, C.c_float(eps), int(relu), st)
            if update and not fused:          # running statistics in another dtype: torch's own arithmetic
                n = N * HW
                with torch.no_grad():
                    if batches is not None:
                        batches.add_(1)
                    f = momentum if momentum >= 0 else 1.0 / float(batches)
                    var = (sums[1::2] / n - (sums[0::2] / n) ** 2).clamp_(min=0).to(running_var.dtype)
                    running_mean.mul_(1 - f).add_(mean.to(running_mean.dtype), alpha=f)
                    running_var.mul_(1 - f).add_(var, alpha=f * n / (n - 1))
        else:
            rm, rv = running_mean.detach().float().contiguous(), running_var.detach().float().contiguous()
            nv.call("hg_bn_apply", nv.ptr(x), nv.ptr(y), None, nv.ptr(rm), nv.ptr(rv), nv.ptr(w), nv.ptr(b), nv.ptr(mean),
                    None, nv.ptr(rstd), N, Cc, HW, C.c_float(eps), int(relu), st)
        ctx.save_for_backward(x, w, b, mean, rstd)
        ctx.cfg = (bool(use_batch_stats), bool(relu), weight is not None, bias is not None)
        # gradient sinks (HyGrid.distributed.FlatGradBucket): dgamma / dbeta are added straight to their bucket slices
        ctx.sinks = (getattr(weight, "_hg_grad_sink", None) if weight is not None else None,
                     getattr(bias, "_hg_grad_sink", None) if bias is not None else None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w, b, mean, rstd = ctx.saved_tensors
        training, relu, has_w, has_b = ctx.cfg
        N, Cc, H, W = x.shape
        HW = H * W
        dev, st = x.device, nv.stream_ptr(x.device)
        dy = dy.contiguous().float()
        dsums = torch.empty(2 * Cc, dtype=torch.float64, device=dev)
        dx = torch.empty_like(x)
        want_g, want_b = has_w and ctx.needs_input_grad[1], has_b and ctx.needs_input_grad[2]
        sink_g, sink_b = ctx.sinks
        # both affine gradients go to the bucket, or neither does (one accumulate flag for the pair)
        sunk = bool(want_g and want_b and sink_g is not None and sink_b is not None
                    and sink_g.view.numel() == Cc and sink_b.view.numel() == Cc)
        if sunk:
            dg, db = sink_g.view, sink_b.view
        else:
            dg = torch.empty(Cc, dtype=torch.float32, device=dev) if want_g else None
            db = torch.empty(Cc, dtype=torch.float32, device=dev) if want_b else None
        nv.call("hg_bn_bwd", nv.ptr(x), nv.ptr(dy), nv.ptr(mean), nv.ptr(rstd), nv.ptr(w), nv.ptr(b), nv.ptr(dsums),
                nv.ptr(dx), nv.ptr(dg), nv.ptr(db), int(sunk), N, Cc, HW, int(relu), int(training), st)
        if sunk:
            sink_g.landed()
            sink_b.landed()
            dg = db = None
        return dx, dg, db, None, None, None, None, None, None, None


def batch_norm_relu(bn: nn.BatchNorm2d, x: torch.Tensor, relu: bool = False) -> torch.Tensor:
    """``relu(bn(x))`` (or ``bn(x)``) with torch.nn.BatchNorm2d's bookkeeping (nn/modules/batchnorm.py ``forward``):
    ``num_batches_tracked`` is incremented and the running averages are blended in training mode when the module tracks
    them -- by the kernels themselves, the step is launch-bound."""
    use_batch_stats = bn.training or (bn.running_mean is None and bn.running_var is None)
    update = bn.training and bn.track_running_stats and bn.running_mean is not None
    momentum = None
    if update:
        momentum = -1.0 if bn.momentum is None else float(bn.momentum)
        if bn.momentum is None and bn.num_batches_tracked is None:
            momentum = 0.0
    needs_running = update or not use_batch_stats
    return _BatchNormReluFn.apply(x, bn.weight, bn.bias, bn.running_mean if needs_running else None,
                                  bn.running_var if needs_running else None, bn.num_batches_tracked if update else None,
                                  use_batch_stats, momentum, bn.eps, relu)
