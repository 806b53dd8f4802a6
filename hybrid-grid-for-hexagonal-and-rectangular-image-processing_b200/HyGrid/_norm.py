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
    def forward(ctx, x, weight, bias, running_mean, running_var, use_batch_stats, factor, eps, relu):
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
            sums = torch.zeros(2 * Cc, dtype=torch.float64, device=dev)
            var = torch.empty(Cc, dtype=torch.float32, device=dev)
            nv.call("hg_bn_stats", nv.ptr(x), nv.ptr(sums), N, Cc, HW, st)
            nv.call("hg_bn_apply", nv.ptr(x), nv.ptr(y), nv.ptr(sums), None, None, nv.ptr(w), nv.ptr(b), nv.ptr(mean),
                    nv.ptr(var), nv.ptr(rstd), N, Cc, HW, C.c_float(eps), int(relu), st)
            if running_mean is not None and factor is not None:
                n = N * HW
                with torch.no_grad():
                    running_mean.mul_(1 - factor).add_(mean, alpha=factor)
                    running_var.mul_(1 - factor).add_(var, alpha=factor * n / (n - 1))
        else:
            rm, rv = running_mean.detach().float().contiguous(), running_var.detach().float().contiguous()
            nv.call("hg_bn_apply", nv.ptr(x), nv.ptr(y), None, nv.ptr(rm), nv.ptr(rv), nv.ptr(w), nv.ptr(b), nv.ptr(mean),
                    None, nv.ptr(rstd), N, Cc, HW, C.c_float(eps), int(relu), st)
        ctx.save_for_backward(x, w, b, mean, rstd)
        ctx.cfg = (bool(use_batch_stats), bool(relu), weight is not None, bias is not None)
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
        dsums = torch.zeros(2 * Cc, dtype=torch.float64, device=dev)
        nv.call("hg_bn_bwd_reduce", nv.ptr(x), nv.ptr(dy), nv.ptr(mean), nv.ptr(rstd), nv.ptr(w), nv.ptr(b), nv.ptr(dsums),
                N, Cc, HW, int(relu), st)
        dx = torch.empty_like(x)
        dg = torch.empty(Cc, dtype=torch.float32, device=dev) if has_w and ctx.needs_input_grad[1] else None
        db = torch.empty(Cc, dtype=torch.float32, device=dev) if has_b and ctx.needs_input_grad[2] else None
        nv.call("hg_bn_bwd_apply", nv.ptr(x), nv.ptr(dy), nv.ptr(mean), nv.ptr(rstd), nv.ptr(w), nv.ptr(b), nv.ptr(dsums),
                nv.ptr(dx), nv.ptr(dg), nv.ptr(db), N, Cc, HW, int(relu), int(training), st)
        return dx, dg, db, None, None, None, None, None, None


def batch_norm_relu(bn: nn.BatchNorm2d, x: torch.Tensor, relu: bool = False) -> torch.Tensor:
    """``relu(bn(x))`` (or ``bn(x)``) with torch.nn.BatchNorm2d's bookkeeping (nn/modules/batchnorm.py ``forward``)."""
    factor = 0.0 if bn.momentum is None else bn.momentum
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        factor = 1.0 / float(bn.num_batches_tracked) if bn.momentum is None else bn.momentum
    use_batch_stats = bn.training or (bn.running_mean is None and bn.running_var is None)
    update = bn.training and bn.track_running_stats
    return _BatchNormReluFn.apply(x, bn.weight, bn.bias, bn.running_mean if (update or not use_batch_stats) else None,
                                  bn.running_var if (update or not use_batch_stats) else None, use_batch_stats,
                                  factor if update else None, bn.eps, relu)
