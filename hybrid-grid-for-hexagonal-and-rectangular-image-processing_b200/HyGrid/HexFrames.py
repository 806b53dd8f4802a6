"""Hex-lattice torch layers -- the API of the reference's ``HyGrid/HexFrames.py`` on sm_100a kernels.

Same names, constructor arguments, public attributes, parameter names (``kernel``, ``bias``) and
error behaviour as the reference (file:line citations are into ``/root/reference/HyGrid/``), but no
doubled "type1" image, no zero-stuffed dense window and no index tables are ever materialised: every
forward / backward is one launch of a hand-written kernel from ``libhygrid_b200.so`` reached through
the C ABI.  CUDA tensors only; there is no CPU path.

Deliberate deviations: ``HexPool2d(stride=None)`` means stride = kernel_size
(the reference crashes, HexFrames.py:272-278); ``HexAdaptivePool2d`` / ``HexGlobalPool2d`` construct
(the reference raises NameError on the undefined ``centroid_pooling``, :360/:408) and only
``method='centroid'`` raises NotImplementedError.
"""
from __future__ import annotations

import ctypes as C
import functools
import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor
from torch.autograd.function import once_differentiable
from torch.nn import init

from . import _native as nv

__all__ = ["pad", "HexConv2d", "HexConv2dAdaptivePadding", "HexPool2d", "HexAdaptivePool2d", "HexGlobalPool2d",
           "heximage_to_type1", "heximage_to_type2", "type1_to_heximage", "max_pooling", "min_pooling",
           "average_pooling", "hexconv2d", "hexpool2d", "HexPixelShuffle", "pixel_shuffle_table",
           "HexConvTranspose2d", "conv_transpose_tables", "set_fp32_tensor_cores"]

_PAD_MODES = {"constant": 0, "reflect": 1, "replicate": 2, "circular": 3}
_POOL = {"max": nv.POOL_MAX, "min": nv.POOL_MIN, "average": nv.POOL_AVG}
_FLOATS = (torch.float32, torch.float64, torch.bfloat16)


def _as4(x: Tensor) -> Tensor:
    while x.dim() < 4:
        x = x.unsqueeze(0)
    return x


# ------------------------------------------------------------------------------------------------
# pad (HexFrames.py:13-21)
# ------------------------------------------------------------------------------------------------
class _Pad2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pl, pr, pt, pb, mode, value):
        x = nv.require_cuda(x, "input").contiguous()
        H, W = x.shape[-2:]
        planes = x.numel() // (H * W)
        y = torch.empty(x.shape[:-2] + (H + pt + pb, W + pl + pr), dtype=x.dtype, device=x.device)
        nv.call("hg_pad2d", nv.ptr(x), nv.ptr(y), planes, H, W, pl, pr, pt, pb, mode, float(value),
                nv.hg_dtype(x.dtype), nv.stream_ptr(x.device))
        ctx.meta = (pl, pr, pt, pb, mode, H, W, planes)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        pl, pr, pt, pb, mode, H, W, planes = ctx.meta
        gy = gy.contiguous()
        gx = torch.empty(gy.shape[:-2] + (H, W), dtype=gy.dtype, device=gy.device)
        nv.call("hg_pad2d_bwd", nv.ptr(gy), nv.ptr(gx), planes, H, W, pl, pr, pt, pb, mode, nv.hg_dtype(gy.dtype),
                nv.stream_ptr(gy.device))
        return gx, None, None, None, None, None, None


def _pad4(x: Tensor, pl: int, pr: int, pt: int, pb: int, mode="constant", value=0) -> Tensor:
    if mode not in _PAD_MODES:
        raise NotImplementedError(f"Unrecognised padding mode {mode}")
    if pl == pr == pt == pb == 0:
        return x
    return _Pad2dFn.apply(x, int(pl), int(pr), int(pt), int(pb), _PAD_MODES[mode], 0 if value is None else value)


def pad(input: Tensor, padding: int = 0, mode="constant", value=0) -> Tensor:
    """``F.pad(input, (padding,)*4, mode, value)`` (HexFrames.py:13-21)."""
    return _pad4(input, padding, padding, padding, padding, mode, value)


# ------------------------------------------------------------------------------------------------
# hex convolution (HexFrames.py:22-185)
# ------------------------------------------------------------------------------------------------
def _conv_desc(x, Cout, Ho, Wo, radius, stride, dilation, groups, pad_, parity, pad_value, y_dtype, algo, relu, pad_mode=0,
               accumulate=0, x_dtype=None):
    N, Cin, H, W = x.shape
    return nv.ConvDesc(N, Cin, Cout, H, W, Ho, Wo, radius, stride, dilation, groups, pad_, parity, float(pad_value),
                       nv.hg_dtype(x_dtype or x.dtype), nv.hg_dtype(y_dtype), algo, int(relu), int(pad_mode) if pad_ else 0,
                       int(accumulate))


# float32 callers on the tensor cores: three bfloat16 passes over split operands (x = x_hi + x_lo, w = w_hi + w_lo;
# y ~ x_hi*w_hi + x_hi*w_lo + x_lo*w_hi, float32 accumulation, the dropped lo*lo term is 2^-16 relative) keep float32-class
# accuracy (measured ~2e-5 of the range; the contract is 1e-4) at a fraction of the CUDA-core stencil's time.
_FP32_TENSOR_CORES = True


def set_fp32_tensor_cores(enabled: bool) -> None:
    """Route float32 (non-autocast) hex convolutions whose channel contraction is dense enough through the tcgen05 kernels
    as three bfloat16 passes over split operands (default on).  Off: the CUDA-core direct stencil, plain float32 FMAs."""
    global _FP32_TENSOR_CORES
    _FP32_TENSOR_CORES = bool(enabled)


def _split_bf16(t: Tensor):
    t = t.contiguous()
    hi = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
    lo = torch.empty_like(hi)
    nv.call("hg_split_bf16", nv.ptr(t), nv.ptr(hi), nv.ptr(lo), t.numel(), nv.stream_ptr(t.device))
    return hi, lo


_eligible_cache = {}


def _x3_eligible(x, Cout, Ho, Wo, radius, stride, dilation, groups, pad_, parity, need_grad):
    """Do the tcgen05 kernels cover all the passes of the split route (forward; data and weight gradient when training)?"""
    if x.shape[1] * Cout < 1024:
        return False
    key = ("x3", tuple(x.shape), Cout, radius, stride, dilation, groups, pad_, parity, bool(need_grad), x.device.index)
    hit = _eligible_cache.get(key)
    if hit is None:
        d = _conv_desc(x, Cout, Ho, Wo, radius, stride, dilation, groups, pad_, parity, 0.0, torch.float32, 2, 0, x_dtype=torch.bfloat16)
        ops = (0, 1, 2) if need_grad else (0,)
        hit = _eligible_cache[key] = all(nv.query("hg_hexconv_umma_eligible", C.byref(d), op) for op in ops)
    return hit


@functools.lru_cache(maxsize=512)
def _conv_out_shape(H, W, radius, stride, dilation, pad_):
    Ho, Wo = C.c_int64(0), C.c_int64(0)
    try:
        nv.call("hg_hexconv_out_shape", H, W, radius, stride, dilation, pad_, C.byref(Ho), C.byref(Wo))
    except nv.HyGridNativeError as e:
        raise ValueError(str(e)) from None
    return Ho.value, Wo.value


def _dense_grad(gy: Tensor, dtype) -> Tensor:
    """The incoming gradient as a dense tensor of ``dtype``.  A broadcast scalar (what ``y.sum().backward()`` hands down:
    every stride 0) is materialised by ``hg_broadcast_fill`` -- 16-byte streaming stores -- instead of torch's strided
    broadcast copy (measured on the C3 stack, 2.1 GB: 1.09 ms for the copy kernel; ``Tensor.fill_(tensor)`` is that same copy)."""
    if gy.numel() and gy.is_cuda and dtype in (torch.float32, torch.bfloat16) and all(st == 0 for st in gy.stride()):
        out = torch.empty(gy.shape, dtype=dtype, device=gy.device)
        scalar = torch.as_strided(gy, (1,), (1,), gy.storage_offset()).to(dtype)
        nv.call("hg_broadcast_fill", nv.ptr(out), nv.ptr(scalar), out.numel(), nv.hg_dtype(dtype), nv.stream_ptr(gy.device))
        return out
    return gy.to(dtype).contiguous()


class _HexConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, bias, meta, scale=None):
        radius, stride, dilation, groups, pad_, parity, pad_value, y_dtype, algo, relu, pad_mode = meta
        x = nv.require_cuda(x, "input").contiguous()
        w = kernel.detach().float().contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        N, Cin, H, W = x.shape
        Cout = w.shape[0]
        Ho, Wo = _conv_out_shape(H, W, radius, stride, dilation, pad_)
        y = torch.empty((N, Cout, Ho, Wo), dtype=y_dtype, device=x.device)
        if algo == 3:              # float32 on the tensor cores: three bfloat16 passes over split operands
            if scale is not None:  # BN-inference affine folded into the weights before the split; b is the shift
                w = w * scale.detach().float().view(-1, 1, 1, 1)
            xh, xl = _split_bf16(x)
            wh, wl = (t.float() for t in _split_bf16(w))
            st = nv.stream_ptr(x.device)
            for k, (xa, wa) in enumerate(((xh, wh), (xh, wl), (xl, wh))):
                dk = _conv_desc(xa, Cout, Ho, Wo, radius, stride, dilation, groups, pad_, parity, pad_value, y_dtype, 2,
                                relu and k == 2, pad_mode, accumulate=int(k > 0))
                nv.call("hg_hexconv_fwd", C.byref(dk), nv.ptr(xa), nv.ptr(wa), nv.ptr(b if k == 0 else None), nv.ptr(y), st)
            if scale is not None:
                ctx.mark_non_differentiable(y)
                return y
            ctx.save_for_backward(xh, xl, wh, wl)
            ctx.meta = meta
            ctx.has_bias = bias is not None
            ctx.param_dtypes = (kernel.dtype, bias.dtype if bias is not None else None)
            ctx.out_shape = (Ho, Wo)
            ctx.x_shape = tuple(x.shape)
            ctx.sinks = (getattr(kernel, "_hg_grad_sink", None), getattr(bias, "_hg_grad_sink", None) if bias is not None else None)
            return y
        d = _conv_desc(x, Cout, Ho, Wo, radius, stride, dilation, groups, pad_, parity, pad_value, y_dtype, algo, relu, pad_mode)
        if scale is not None:      # inference-only fused per-channel affine (HexConvModule conv -> BN(eval) -> ReLU); b is the shift
            sc = scale.detach().float().contiguous()
            nv.call("hg_hexconv_fwd_affine", C.byref(d), nv.ptr(x), nv.ptr(w), nv.ptr(sc), nv.ptr(b), nv.ptr(y),
                    nv.stream_ptr(x.device))
            ctx.mark_non_differentiable(y)
            return y
        nv.call("hg_hexconv_fwd", C.byref(d), nv.ptr(x), nv.ptr(w), nv.ptr(b), nv.ptr(y), nv.stream_ptr(x.device))
        ctx.save_for_backward(x, w)
        ctx.meta = meta
        ctx.has_bias = bias is not None
        ctx.param_dtypes = (kernel.dtype, bias.dtype if bias is not None else None)
        ctx.out_shape = (Ho, Wo)
        # gradient sinks (HyGrid.distributed.FlatGradBucket): the weight-gradient kernel accumulates straight into the
        # all-reduce bucket instead of a zero-filled temporary that autograd then adds to .grad
        ctx.sinks = (getattr(kernel, "_hg_grad_sink", None), getattr(bias, "_hg_grad_sink", None) if bias is not None else None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        if ctx.meta[8] == 3:
            return _HexConvFn._backward_x3(ctx, gy)
        x, w = ctx.saved_tensors
        radius, stride, dilation, groups, pad_, parity, pad_value, y_dtype, algo, relu, pad_mode = ctx.meta
        if relu:
            raise RuntimeError("the fused ReLU epilogue is inference-only")
        Ho, Wo = ctx.out_shape
        gy = _dense_grad(gy, y_dtype)
        d = _conv_desc(x, w.shape[0], Ho, Wo, radius, stride, dilation, groups, pad_, parity, pad_value, y_dtype, algo, 0, pad_mode)
        st = nv.stream_ptr(x.device)
        framed = bool(pad_) and pad_mode != 0          # reflect / replicate / circular frame, resolved inside the loaders

        def pick(op, desc=None):   # a forced tcgen05 forward does not oblige the backward ops to have a tcgen05 kernel
            desc = d if desc is None else desc
            desc.algo = algo if (algo != 2 or nv.query("hg_hexconv_umma_eligible", C.byref(desc), op)) else 1
            return desc
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if framed:
                # the data gradient of the frame folds back onto the image: gradient of the virtually padded input
                # (pad = 0 on the [H + 2p, W + 2p] geometry), then the pad kernel's adjoint
                N, Cin, H, W = x.shape
                gp = torch.empty((N, Cin, H + 2 * pad_, W + 2 * pad_), dtype=x.dtype, device=x.device)
                dp = _conv_desc(gp, w.shape[0], Ho, Wo, radius, stride, dilation, groups, 0, parity, 0.0, y_dtype, algo, 0)
                nv.call("hg_hexconv_dgrad", C.byref(pick(1, dp)), nv.ptr(gy), nv.ptr(w), nv.ptr(gp), st)
                gx = torch.empty_like(x)
                nv.call("hg_pad2d_bwd", nv.ptr(gp), nv.ptr(gx), N * Cin, H, W, pad_, pad_, pad_, pad_, pad_mode, nv.hg_dtype(x.dtype), st)
            else:
                gx = torch.empty_like(x)
                nv.call("hg_hexconv_dgrad", C.byref(pick(1)), nv.ptr(gy), nv.ptr(w), nv.ptr(gx), st)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw = pick(2)
            if x.shape[1] < 16 and groups == 1 and algo != 1 and x.dtype == torch.bfloat16:
                # first layers (RGB under autocast): the tcgen05 weight-gradient kernel takes them with the channel slice
                # rounded up to 16 inside the kernel (zeros for the channels that do not exist); 20x faster than the
                # CUDA-core stencil, and no zero-padded copy of x
                forced = _conv_desc(x, w.shape[0], Ho, Wo, radius, stride, dilation, groups, pad_, parity, pad_value, y_dtype, 2, 0, pad_mode)
                if nv.query("hg_hexconv_umma_eligible", C.byref(forced), 2):
                    dw = forced
            sink_w, sink_b = ctx.sinks
            if sink_w is None or tuple(sink_w.view.shape) != tuple(w.shape):
                sink_w = None
            if not ctx.has_bias or sink_b is None:
                sink_b = None
            gw = sink_w.view if sink_w is not None else torch.zeros(w.shape, dtype=torch.float32, device=x.device)
            gb = None
            if ctx.has_bias:
                gb = sink_b.view if sink_b is not None else torch.zeros(w.shape[0], dtype=torch.float32, device=x.device)
            nv.call("hg_hexconv_wgrad", C.byref(dw), nv.ptr(x), nv.ptr(gy), nv.ptr(gw), nv.ptr(gb), st)
            if sink_w is not None:
                gw = None
                sink_w.landed()
            else:
                gw = gw.to(ctx.param_dtypes[0])
            if sink_b is not None:
                gb = None
                sink_b.landed()
            elif gb is not None:
                gb = gb.to(ctx.param_dtypes[1])
        return gx, gw, gb, None, None


def _backward_x3(ctx, gy):
    """Backward of the split float32 route: data gradient = three accumulating tcgen05 passes over (gy_hi, gy_lo) x (w_hi, w_lo),
    weight gradient = three accumulating passes over (x_hi, x_lo) x (gy_hi, gy_lo); the bias gradient sums gy_hi + gy_lo."""
    xh, xl, wh, wl = ctx.saved_tensors
    radius, stride, dilation, groups, pad_, parity, pad_value, y_dtype, algo, relu, pad_mode = ctx.meta
    if relu:
        raise RuntimeError("the fused ReLU epilogue is inference-only")
    Ho, Wo = ctx.out_shape
    N, Cin, H, W = ctx.x_shape
    Cout = wh.shape[0]
    gh, gl = _split_bf16(_dense_grad(gy, torch.float32))
    st = nv.stream_ptr(xh.device)
    bf = torch.bfloat16
    framed = bool(pad_) and pad_mode != 0
    gx = gw = gb = None
    if ctx.needs_input_grad[0]:
        if framed:      # gradient of the virtually padded input, then the pad kernel's adjoint (see the plain backward)
            tgt = torch.empty((N, Cin, H + 2 * pad_, W + 2 * pad_), dtype=torch.float32, device=xh.device)
            geo = (tgt, 0, 0.0, 0)
        else:
            tgt = torch.empty((N, Cin, H, W), dtype=torch.float32, device=xh.device)
            geo = (tgt, pad_, pad_value, pad_mode)
        for k, (ga, wa) in enumerate(((gh, wh), (gh, wl), (gl, wh))):
            dk = _conv_desc(geo[0], Cout, Ho, Wo, radius, stride, dilation, groups, geo[1], parity, geo[2], bf, 2, 0, geo[3],
                            accumulate=int(k > 0))
            nv.call("hg_hexconv_dgrad", C.byref(dk), nv.ptr(ga), nv.ptr(wa), nv.ptr(tgt), st)
        if framed:
            gx = torch.empty((N, Cin, H, W), dtype=torch.float32, device=xh.device)
            nv.call("hg_pad2d_bwd", nv.ptr(tgt), nv.ptr(gx), N * Cin, H, W, pad_, pad_, pad_, pad_, pad_mode, nv.F32, st)
        else:
            gx = tgt
    if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
        sink_w, sink_b = ctx.sinks
        if sink_w is None or tuple(sink_w.view.shape) != tuple(wh.shape):
            sink_w = None
        if not ctx.has_bias or sink_b is None:
            sink_b = None
        gw = sink_w.view if sink_w is not None else torch.zeros(wh.shape, dtype=torch.float32, device=xh.device)
        if ctx.has_bias:
            gb = sink_b.view if sink_b is not None else torch.zeros(Cout, dtype=torch.float32, device=xh.device)
        dw = _conv_desc(xh, Cout, Ho, Wo, radius, stride, dilation, groups, pad_, parity, pad_value, bf, 2, 0, pad_mode)
        for k, (xa, ga) in enumerate(((xh, gh), (xh, gl), (xl, gh))):
            nv.call("hg_hexconv_wgrad", C.byref(dw), nv.ptr(xa), nv.ptr(ga), nv.ptr(gw), nv.ptr(gb if k < 2 else None), st)
        if sink_w is not None:
            gw = None
            sink_w.landed()
        else:
            gw = gw.to(ctx.param_dtypes[0])
        if sink_b is not None:
            gb = None
            sink_b.landed()
        elif gb is not None:
            gb = gb.to(ctx.param_dtypes[1])
    return gx, gw, gb, None, None


_HexConvFn._backward_x3 = staticmethod(_backward_x3)


def hexconv2d(x: Tensor, kernel: Tensor, bias: Optional[Tensor] = None, even_odd_offset=0, radius=2, stride=1, padding=0,
              dilation=1, groups=1, padding_value=0.0, out_dtype=torch.float32, algo=0, relu=False, padding_mode='constant') -> Tensor:
    """Functional hex convolution with virtual padding in any F.pad mode (closed form of HexFrames.py:96-169)."""
    x = _as4(x)
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"hex convolution runs on float32 or bfloat16 activations, got {x.dtype}")
    parity = (int(even_odd_offset) + int(padding)) % 2
    if padding_mode not in _PAD_MODES:
        raise NotImplementedError(f"Unrecognised padding mode {padding_mode}")
    meta = (int(radius), int(stride), int(dilation), int(groups), int(padding), parity, float(padding_value or 0),
            out_dtype, int(algo), bool(relu), _PAD_MODES[padding_mode])
    return _HexConvFn.apply(x, kernel, bias, meta)


class HexConv2d(nn.Module):
    """Hexagonal-footprint convolution on an offset-stored hex lattice (HexFrames.py:22-185).

    Parameters ``kernel`` ``[out, in/groups, 1, 3r^2-3r+1]`` and ``bias`` ``[out]`` keep the reference's
    names, shapes and initialisation (:74-95), so ``state_dict``s interchange.  The output is float32
    (the reference interleaves into a float32 buffer, :157-160); under ``torch.autocast(bfloat16)`` the
    activations are read as bfloat16 with float32 accumulation."""

    def __init__(self, in_channels, out_channels, even_odd_offset, hexkernel_radius, stride=1,
                 padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='constant', padding_value=0):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.even_odd_offset = even_odd_offset
        self.padded_even_odd_offset = (even_odd_offset + padding) % 2
        self.hexkernel_radius = hexkernel_radius
        self.hexkernel_size = 2 * hexkernel_radius - 1
        self.kernelnum = 3 * hexkernel_radius ** 2 - 3 * hexkernel_radius + 1
        self.stride = stride
        self.sh = stride
        self.sw = stride * 2
        self.out_even_odd_offset = 0
        self.pad = padding
        self.groups = groups
        self.b = bias
        self.dilation = dilation
        self.padding_mode = padding_mode
        self.padding_value = padding_value
        if in_channels % groups != 0:
            raise ValueError('in_channels must be divisible by groups')
        if out_channels % groups != 0:
            raise ValueError('out_channels must be divisible by groups')
        self.kernel = nn.Parameter(torch.empty([out_channels, in_channels // groups, 1, self.kernelnum], dtype=torch.float))
        if self.b == True:  # noqa: E712  (the reference's own test)
            self.bias = nn.Parameter(torch.empty([out_channels, ]))
        else:
            self.register_parameter('bias', None)
        self.k_w = 2 * self.dilation * (2 * self.hexkernel_radius - 2) + 1
        self.k_h = (self.hexkernel_size - 1) * self.dilation + 1
        self.algo = 0            # 0 auto, 1 direct stencil, 2 tcgen05 implicit GEMM
        self.out_dtype = torch.float32
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.kernel, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.kernel)
            if fan_in != 0:
                bound = 1 / math.sqrt(fan_in)
                init.uniform_(self.bias, -bound, bound)

    def _tensor_core_ok(self, input: Tensor, pad_=None) -> bool:
        """Would the tcgen05 kernel take this layer?  Then fp32 activations are fed to it directly (it rounds
        them to bfloat16 on the way into shared memory) instead of paying a separate cast pass."""
        x = _as4(input)
        if x.dtype not in (torch.float32, torch.bfloat16) or x.dim() != 4:
            return False
        pad_ = self.pad if pad_ is None else pad_
        key = ("tc", tuple(x.shape), x.dtype, self.out_channels, self.hexkernel_radius, self.stride, self.dilation, self.groups, pad_,
               self.padded_even_odd_offset, self.out_dtype, x.device.index)
        hit = _eligible_cache.get(key)
        if hit is None:
            try:
                Ho, Wo = _conv_out_shape(x.shape[2], x.shape[3], self.hexkernel_radius, self.stride, self.dilation, pad_)
                d = _conv_desc(x, self.out_channels, Ho, Wo, self.hexkernel_radius, self.stride, self.dilation, self.groups,
                               pad_, self.padded_even_odd_offset, 0.0, self.out_dtype, 2, 0)
                hit = bool(nv.query("hg_hexconv_umma_eligible", C.byref(d), 0))
            except ValueError:
                hit = False
            _eligible_cache[key] = hit
        return hit

    def _activation(self, input: Tensor, pad_=None):
        """(input in the dtype the kernels read, whether autocast routes this call to the tcgen05 kernel)."""
        if torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            if self.algo != 1 and self._tensor_core_ok(input, pad_):
                return (input if input.dtype in (torch.float32, torch.bfloat16) else input.float()), True
            return input.to(torch.bfloat16), False
        if self.kernel.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"HexConv2d parameters must be float32 (or bfloat16), got {self.kernel.dtype}")
        return input.to(self.kernel.dtype), False

    def forward(self, input: Tensor, relu: bool = False, affine=None, frame=None) -> Tensor:
        """``affine=(scale, shift)``: inference-only fused ``act(conv(x) * scale[c] + shift[c])`` (the conv's own bias must
        already be folded into ``shift``); ``relu=True`` fuses the ReLU.  Neither records anything for backward.
        ``frame=(p, mode)``: an explicit padding layer in front of a ``padding=0`` conv (HexConvModule, HexModules.py:185-190)
        folded into the kernel -- ``p`` cells of ``mode`` around the input, WITHOUT the parity change ``padding=p`` would
        make (the reference builds that conv with padding 0, so its row parity ignores the frame)."""
        input, autocast_tc = self._activation(input, None if frame is None else int(frame[0]))
        self._autocast_tc = autocast_tc          # record of the last call (introspection / tests); not read by the forward
        input = _as4(input)
        algo = 2 if autocast_tc else self.algo
        pad_, parity, mode, value = self.pad, self.padded_even_odd_offset, self.padding_mode, self.padding_value
        if frame is not None:
            if self.pad:
                raise ValueError("frame= is for convolutions built with padding=0")
            pad_, mode, value = int(frame[0]), frame[1], 0
        if mode not in _PAD_MODES:
            raise NotImplementedError(f"Unrecognised padding mode {mode}")
        if pad_ and mode != 'constant':          # the loaders read the reflected / replicated / wrapped image in place
            H, W = input.shape[-2:]
            if (mode == 'reflect' and (pad_ >= H or pad_ >= W)) or (mode == 'circular' and (pad_ > H or pad_ > W)):
                raise RuntimeError(f"{mode} padding of {pad_} does not fit a {H} x {W} input")
        if (algo == 0 and _FP32_TENSOR_CORES and input.dtype == torch.float32 and self.out_dtype == torch.float32
                and self.kernel.dtype == torch.float32 and input.is_cuda and not autocast_tc):
            try:
                Ho, Wo = _conv_out_shape(input.shape[2], input.shape[3], self.hexkernel_radius, self.stride, self.dilation, pad_)
                need_grad = torch.is_grad_enabled() and (input.requires_grad or self.kernel.requires_grad) and affine is None
                if _x3_eligible(input, self.out_channels, Ho, Wo, self.hexkernel_radius, self.stride, self.dilation, self.groups,
                                pad_, parity, need_grad):
                    algo = 3
            except ValueError:
                pass
        meta = (self.hexkernel_radius, self.stride, self.dilation, self.groups, pad_, parity,
                float(value or 0), self.out_dtype, algo, bool(relu), _PAD_MODES[mode])
        if affine is not None:
            if torch.is_grad_enabled() and (input.requires_grad or self.kernel.requires_grad):
                raise RuntimeError("the fused affine epilogue is inference-only: call it under torch.no_grad()")
            return _HexConvFn.apply(input, self.kernel, affine[1], meta, affine[0])
        return _HexConvFn.apply(input, self.kernel, self.bias, meta)

    def extra_repr(self):
        s = ('{in_channels}, {out_channels}, kernel_radius={hexkernel_radius}'
             ', stride={stride}')
        if self.pad != (0,):
            s += ', padding={pad}'
        if self.dilation != (1,):
            s += ', dilation={dilation}'
        if self.groups != 1:
            s += ', groups={groups}'
        if self.bias is None:
            s += ', bias=False'
        if self.padding_mode != 'zeros':
            s += ', padding_mode={padding_mode}'
        return s.format(**self.__dict__)


class HexConv2dAdaptivePadding(HexConv2d):
    """TF-"same" variant (HexFrames.py:187-253): ignores ``padding`` and pads so that the filter covers
    the whole input.  Like the reference, the parity flip caused by the top pad is not compensated."""

    def __init__(self, in_channels: int, out_channels: int, even_odd_offset: int, hexkernel_radius: int,
                 stride: int = 1, padding: int = 0, dilation: int = 1, groups: int = 1, bias: bool = True):
        super().__init__(in_channels, out_channels, even_odd_offset=even_odd_offset,
                         hexkernel_radius=hexkernel_radius, stride=stride, padding=0, dilation=dilation,
                         groups=groups, bias=bias)

    def __repr__(self):
        return (f"HexConv2dAdaptivePadding({self.in_channels}, {self.out_channels}, kernel_radius={self.hexkernel_radius}, "
                f"stride={self.sh}, padding={self.pad}, dilation={self.dilation}, groups={self.groups}, bias={self.b})")

    def forward(self, input: Tensor, relu: bool = False, affine=None) -> Tensor:
        input = pad(input, self.pad, self.padding_mode, self.padding_value)
        self.pad = 0
        img_h, img_w = input.size()[-2:]
        kernel_size = self.hexkernel_radius * 2 - 1
        stride = self.stride
        output_h = math.ceil(img_h / stride)
        output_w = math.ceil(img_w / stride)
        pad_h = max((output_h - 1) * self.stride + (kernel_size - 1) * self.dilation + 1 - img_h, 0)
        pad_w = max(output_w * self.stride + (kernel_size - 1) * self.dilation + 1 - img_w, 0)
        input = _as4(self._activation(input)[0])
        if pad_h > 0 or pad_w > 0:
            input = _pad4(input, pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2)
        return super().forward(input, relu, affine)


# ------------------------------------------------------------------------------------------------
# pooling (HexFrames.py:255-414, reductions :461-479)
# ------------------------------------------------------------------------------------------------
class _HexPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, method, geom):
        kh, kw, sh, sw, shift, pad_, pad_value, tail_h, tail_w, tail_value, hn, wn = geom
        x = nv.require_cuda(x, "input").contiguous()
        B, Cc, H, W = x.shape
        planes = B * Cc
        y = torch.empty((B, Cc, hn, wn), dtype=x.dtype, device=x.device)
        need_aux = ctx.needs_input_grad[0]
        aux_bytes = 1 if kh * kw <= 127 else 4
        aux = (torch.empty((B, Cc, hn, wn), dtype=torch.int8 if aux_bytes == 1 else torch.int32, device=x.device)
               if need_aux else None)
        try:
            nv.call("hg_hexpool_fwd", nv.ptr(x), nv.ptr(y), nv.ptr(aux), aux_bytes, planes, H, W, hn, wn, kh, kw, sh, sw,
                    shift, pad_, float(pad_value), tail_h, tail_w, float(tail_value), method, nv.hg_dtype(x.dtype),
                    nv.stream_ptr(x.device))
        except nv.HyGridNativeError as e:
            if "leaves the image" in str(e):
                raise IndexError(str(e)) from None
            raise
        ctx.geom, ctx.method, ctx.aux_bytes, ctx.in_hw = geom, method, aux_bytes, (H, W)
        if need_aux:
            ctx.save_for_backward(aux, x if method == nv.POOL_AVG else None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        aux, x = ctx.saved_tensors
        kh, kw, sh, sw, shift, pad_, _, _, _, _, hn, wn = ctx.geom
        gy = gy.contiguous()
        B, Cc = gy.shape[:2]
        H, W = ctx.in_hw
        gx = torch.empty((B, Cc, H, W), dtype=gy.dtype, device=gy.device)
        nv.call("hg_hexpool_bwd", nv.ptr(gy), nv.ptr(aux), ctx.aux_bytes, nv.ptr(x), nv.ptr(gx), B * Cc, H, W, hn, wn,
                kh, kw, sh, sw, shift, pad_, ctx.method, nv.hg_dtype(gy.dtype), nv.stream_ptr(gy.device))
        return gx, None, None


def _pool_apply(x, method, geom):
    x = _as4(x)
    if x.dtype not in _FLOATS:
        raise TypeError(f"hex pooling runs on float32 / float64 / bfloat16 tensors, got {x.dtype}")
    return _HexPoolFn.apply(x, method, geom)


def hexpool2d(x: Tensor, method: str, kh: int, kw: int, sh: int, sw: int, shift: int, padding=0, padding_value=0.0,
              tail_h=0, tail_w=0, tail_value=0.0) -> Tensor:
    """Windows ``rows sh*I + a, cols ((I%2)*shift)//2 + J*sw + b`` over the virtually padded input."""
    x = _as4(x)
    H, W = x.shape[-2:]
    h, w = H + 2 * padding + tail_h, W + 2 * padding + tail_w
    hn = (h - kh) // sh + 1
    wn = (w - sw // 2) // sw
    geom = (kh, kw, sh, sw, shift, padding, padding_value, tail_h, tail_w, tail_value, max(hn, 0), max(wn, 0))
    return _pool_apply(x, _POOL[method], geom)


def centroid_pooling(input):
    raise NotImplementedError("'centroid' pooling is referenced but never defined in the reference "
                              "(HexFrames.py:360, :408)")


class HexPool2d(nn.Module):
    """Hex pooling (HexFrames.py:255-341): odd output rows start half a stride to the right;
    ``even_odd_offset`` is stored but, like in the reference, does not enter the window arithmetic."""

    def __init__(self, method, kernel_size=2, stride=None, padding=0, even_odd_offset=0,
                 padding_mode='constant', padding_value=0, ceil_mode: bool = False,
                 count_include_pad: bool = True, divisor_override: Optional[int] = None):
        super().__init__()
        self.out_offset = 0
        self.offset = (even_odd_offset + padding) % 2
        self.PoolingMethods = {'max': max_pooling, 'min': min_pooling, 'average': average_pooling}
        self.method = self.PoolingMethods[method]
        self._method_name = method
        if isinstance(kernel_size, int):
            kernel_size = [kernel_size, kernel_size]
        self.kernel_size = kernel_size
        self.kh, self.kw = kernel_size
        if stride is None:          # deviation: the reference overwrites this with None and crashes
            stride = list(kernel_size)
        if isinstance(stride, int):
            stride = [stride, stride]
        self.stride = stride
        self.sh, self.sw = self.stride
        self.padding = padding
        self.padding_mode = padding_mode
        self.padding_value = padding_value
        self.ceil_mode = ceil_mode
        self.count_include_pad = count_include_pad

    def forward(self, input):
        input = _as4(input)
        pad_ = self.padding
        if pad_ and self.padding_mode != 'constant':
            input = pad(input, pad_, self.padding_mode, self.padding_value)
            pad_ = 0
        H, W = input.shape[-2:]
        h, w = H + 2 * pad_, W + 2 * pad_
        self.hn = h // self.sh
        self.wn = (w - self.sw // 2 - self.sw) // self.sw + 1
        tail_h = tail_w = 0
        tail_value = 0.0
        if self.ceil_mode:
            ph = (self.kh - h + self.hn * self.sh) % self.kh
            pw = (self.kw - w + (self.wn * self.sw + self.sw // 2)) % self.kw
            # literal argument order of HexFrames.py:297-298: F.pad(input, (0, ph, 0, pw)) pads ph COLUMNS and pw ROWS
            tail_w, tail_h = ph, pw
            tail_value = 0.0 if self.count_include_pad else float('nan')
        h, w = h + tail_h, w + tail_w
        self.hn = (h - self.kh) // self.sh + 1
        self.wn = (w - self.sw // 2) // self.sw
        geom = (self.kh, self.kw, self.sh, self.sw, self.sw, pad_, float(self.padding_value or 0), tail_h, tail_w,
                tail_value, max(self.hn, 0), max(self.wn, 0))
        return _pool_apply(input, _POOL[self._method_name], geom)

    def extra_repr(self) -> str:
        return 'kernel_size={}, stride={}, padding={}'.format(self.kernel_size, self.stride, self.padding)


class HexAdaptivePool2d(nn.Module):
    """HexFrames.py:344-401: ``outsize`` must be an int; windows of ``int(h/n) x int(w/(n+.5))``."""

    def __init__(self, outsize, method, padding=0, padding_mode='constant', padding_value=0):
        super().__init__()
        if isinstance(outsize, int):
            outsize = [outsize, outsize]
        else:
            raise Exception('outsize = 整数 s 或者列表[h, w]，其它的不行')
        self.hn, self.wn = outsize
        self.PoolingMethods = {'max': max_pooling, 'min': min_pooling, 'average': average_pooling,
                               'centroid': centroid_pooling}
        self.method = self.PoolingMethods[method]
        self._method_name = method

    def forward(self, input):
        if self._method_name == 'centroid':
            centroid_pooling(input)
        input = _as4(input)
        h, w = input.shape[-2:]
        grid_h = int(h / self.hn)
        grid_w = int(w / (self.wn + 0.5)) if grid_h > 1 else int(w / self.wn)
        if grid_h < 1 or grid_w < 1:
            raise IndexError("adaptive pool window is empty")
        geom = (grid_h, grid_w, grid_h, grid_w, grid_w, 0, 0.0, 0, 0, 0.0, self.hn, self.wn)
        return _pool_apply(input, _POOL[self._method_name], geom)


class _ReduceLastFn(torch.autograd.Function):
    """NaN-aware max / min / mean over the last dimension (HexFrames.py:461-479)."""

    @staticmethod
    def forward(ctx, x, method):
        x = nv.require_cuda(x, "input").contiguous()
        if x.dtype not in _FLOATS:
            raise TypeError(f"hex pooling runs on float32 / float64 / bfloat16 tensors, got {x.dtype}")
        L = x.shape[-1]
        planes = x.numel() // L if L else 0
        y = torch.empty(x.shape[:-1], dtype=x.dtype, device=x.device)
        aux = torch.empty(x.shape[:-1], dtype=torch.int32, device=x.device)
        nv.call("hg_hexglobalpool_fwd", nv.ptr(x), nv.ptr(y), nv.ptr(aux), planes, L, method, nv.hg_dtype(x.dtype),
                nv.stream_ptr(x.device))
        ctx.save_for_backward(x, aux)
        ctx.method = method
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, aux = ctx.saved_tensors
        gy = gy.contiguous()
        L = x.shape[-1]
        gx = torch.empty_like(x)
        nv.call("hg_hexglobalpool_bwd", nv.ptr(gy), nv.ptr(x), nv.ptr(aux), nv.ptr(gx), x.numel() // L, L, ctx.method,
                nv.hg_dtype(x.dtype), nv.stream_ptr(x.device))
        return gx, None


def max_pooling(input):
    return _ReduceLastFn.apply(input, nv.POOL_MAX)


def min_pooling(input):
    return _ReduceLastFn.apply(input, nv.POOL_MIN)


def average_pooling(input):
    return _ReduceLastFn.apply(input, nv.POOL_AVG)


class HexGlobalPool2d(nn.Module):
    """HexFrames.py:402-414: reduce over H*W, returns ``(B, C)`` (no trailing 1x1)."""

    def __init__(self, method):
        super().__init__()
        self.PoolingMethods = {'max': max_pooling, 'min': min_pooling, 'average': average_pooling,
                               'centroid': centroid_pooling}
        self.method = self.PoolingMethods[method]

    def forward(self, input):
        input = _as4(input)
        unfolded = input.reshape(input.size(0), input.size(1), input.size(2) * input.size(3))
        return self.method(unfolded)


# ------------------------------------------------------------------------------------------------
# format conversion (HexFrames.py:417-458)
# ------------------------------------------------------------------------------------------------
class _ToTypeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, offset, rows_mul):
        x = nv.require_cuda(x, "input").contiguous()
        B, Cc, H, W = x.shape
        y = torch.empty((B, Cc, rows_mul * H, 2 * W + 1), dtype=torch.float32, device=x.device)
        nv.call("hg_hex_to_type1" if rows_mul == 1 else "hg_hex_to_type2", nv.ptr(x), nv.ptr(y), B * Cc, H, W, offset % 2,
                nv.hg_dtype(x.dtype), nv.F32, nv.stream_ptr(x.device))
        ctx.meta = (offset % 2, rows_mul, x.dtype)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        # every cell was written 2 (type1) or 4 (type2) times: the gradient is the sum of its copies.
        # Rarely needed (the converters sit outside the conv path here), so composed from strided views.
        offset, rows_mul, dtype = ctx.meta
        g = gy if rows_mul == 1 else gy[:, :, 0::2] + gy[:, :, 1::2]
        H = g.shape[2]
        W = (g.shape[3] - 1) // 2
        gx = torch.empty(g.shape[:2] + (H, W), dtype=g.dtype, device=g.device)
        for par in (0, 1):
            s = (par + offset) % 2
            rows = g[:, :, par::2]
            gx[:, :, par::2] = rows[..., s:s + 2 * W:2] + rows[..., s + 1:s + 1 + 2 * W:2]
        return gx.to(dtype), None, None


def heximage_to_type1(input: Tensor, even_odd_offset) -> Tensor:
    """(B,C,H,W) -> (B,C,H,2W+1) float32 doubled raster (HexFrames.py:417-445)."""
    return _ToTypeFn.apply(_as4(input), int(even_odd_offset), 1)


def heximage_to_type2(input: Tensor, even_odd_offset) -> Tensor:
    """type1 with every row repeated (HexFrames.py:446-449)."""
    return _ToTypeFn.apply(_as4(input), int(even_odd_offset), 2)


def type1_to_heximage(input: Tensor, even_odd_offset: int):
    """``input[:, :, :, 1::2]`` (HexFrames.py:450-458): a strided view, exactly like the reference."""
    out = input[:, :, :, 1::2]
    return out, even_odd_offset


# ------------------------------------------------------------------------------------------------
# hex pixel shuffle (retired from the reference: "codes in old versions.txt":68-126)
# ------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=64)
def pixel_shuffle_table(r: int, H: int, W: int):
    """Which input element lands in every output cell of ``HexPixelShuffle(r)`` on an ``H x W`` lattice:
    int64 arrays ``(n, row, col)`` of shape ``(Ho, Wo)`` -- sub-channel block, input row, input column -- ``n = -1``
    where nothing lands (zero).  Closed form of the reference's slice assignments (old versions :95-126):

    * the doubled canvas has ``r*H + r - 1`` rows and ``r*W + r//2`` hex cells per row; the result is the canvas
      without its first / last ``r - 1`` rows and ``r//2`` / ``ceil(r/2)`` cells (:85-87, :126);
    * sub-channel ``n`` enumerates the cells of a radius-``r`` hexagon row by row (row ``i`` of ``2r-1`` holds
      ``r - t`` cells, ``t = |1 + i - r|``, :101-104, :123);
    * even input rows ``2a`` write canvas row ``i + 2r*a``, sub-column ``1 + t + 2k + 2r*b`` and its neighbour at
      ``+1`` (``r`` even) or ``-1`` (``r`` odd); odd input rows ``2a + 1`` the same ``r`` rows and ``r`` sub-columns
      further (:105-122); only odd sub-columns survive ``type1_to_heximage`` (``[..., 1::2]``, :125);
    * where two writes meet, the later one (larger ``n``, odd rows after even rows) stays."""
    Hc, Wc = r * H + r - 1, r * W + r // 2
    Yc, Xc = np.meshgrid(np.arange(r - 1, Hc - (r - 1)), np.arange(r // 2, Wc - (r + 1) // 2), indexing="ij")
    starts = np.cumsum([0] + [r - abs(1 + i - r) for i in range(2 * r - 1)])
    n_out = np.full(Yc.shape, -1, np.int64)
    row_out, col_out, order = np.zeros_like(n_out), np.zeros_like(n_out), np.full(Yc.shape, -1, np.int64)
    pair = 1 if r % 2 == 0 else -1
    for par in (0, 1):
        i = (Yc - par * r) % (2 * r)
        ok = (i <= 2 * r - 2) & (Yc - par * r - i >= 0)
        i = np.minimum(i, 2 * r - 2)
        t = np.abs(1 + i - r)
        first = 1 + t + par * r                                   # canvas sub-column of k = 0, b = 0
        base = (np.where(first % 2 == 1, first, first + pair) - 1) // 2
        d = Xc - base
        k, b = d % r, d // r
        row = 2 * ((Yc - par * r - i) // (2 * r)) + par
        ok &= (d >= 0) & (k < r - t) & (b < W) & (row < H)
        n = starts[i] + k
        take = ok & (2 * n + par > order)
        n_out[take], row_out[take], col_out[take], order[take] = n[take], row[take], b[take], (2 * n + par)[take]
    return n_out, row_out, col_out


@functools.lru_cache(maxsize=64)
def _pixel_shuffle_offsets(r, H, W, cout, device):
    n, row, col = pixel_shuffle_table(r, H, W)
    off = np.where(n >= 0, n * (cout * H * W) + row * W + col, -1).astype(np.int64)
    if np.unique(off[off >= 0]).size != int((off >= 0).sum()):
        raise RuntimeError("hex pixel shuffle table is not injective")       # the scatter adjoint relies on it
    return torch.from_numpy(np.ascontiguousarray(off.reshape(-1))).to(torch.device(device)), off.shape


def _injective_layers(off: np.ndarray):
    """Split a gather table into tables that each read every source element at most once (the j-th reader of an
    element goes to layer j; other entries are -1).  The adjoint of the gather is then one plain scatter per layer,
    summed -- no atomics, deterministic.  A table that is already injective comes back as itself."""
    off = np.ascontiguousarray(off.reshape(-1))
    valid = np.flatnonzero(off >= 0)
    order = valid[np.argsort(off[valid], kind="stable")]
    sorted_off = off[order]
    first = np.r_[True, sorted_off[1:] != sorted_off[:-1]] if order.size else np.zeros(0, bool)
    start = np.maximum.accumulate(np.where(first, np.arange(order.size), 0)) if order.size else np.zeros(0, np.int64)
    rank = np.arange(order.size) - start                     # 0 for the first reader of an element, 1 for the second, ...
    layers = []
    for j in range(int(rank.max()) + 1 if order.size else 1):
        t = np.full_like(off, -1)
        pick = order[rank == j]
        t[pick] = off[pick]
        layers.append(t)
    return layers


class _PlaneGatherFn(torch.autograd.Function):
    """``y[b, c, e] = x[b].flatten()[c*chan_stride + table[e]]`` (0 where ``table[e] < 0``), float32 out.
    ``bwd_tables``: the table split into injective layers (``_injective_layers``) for the scatter adjoint."""

    @staticmethod
    def forward(ctx, x, table, out_hw, chans, chan_stride, bwd_tables=None):
        x = nv.require_cuda(x, "input").contiguous()
        B = x.shape[0]
        batch_stride = x.numel() // B if B else 0
        y = torch.empty((B, chans) + tuple(out_hw), dtype=torch.float32, device=x.device)
        nv.call("hg_plane_gather", nv.ptr(x), nv.ptr(y), nv.ptr(table), B, chans, out_hw[0] * out_hw[1], batch_stride,
                chan_stride, nv.hg_dtype(x.dtype), nv.F32, nv.stream_ptr(x.device))
        ctx.save_for_backward(*(bwd_tables if bwd_tables is not None else (table,)))
        ctx.meta = (tuple(x.shape), x.dtype, chans, chan_stride, batch_stride, tuple(out_hw))
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        shape, dtype, chans, chan_stride, batch_stride, out_hw = ctx.meta
        gy = gy.contiguous().float()
        gx = None
        for layer in ctx.saved_tensors:
            part = torch.zeros(shape, dtype=torch.float32, device=gy.device)
            nv.call("hg_plane_scatter", nv.ptr(gy), nv.ptr(part), nv.ptr(layer), shape[0], chans, out_hw[0] * out_hw[1],
                    batch_stride, chan_stride, nv.F32, nv.stream_ptr(gy.device))
            gx = part if gx is None else gx.add_(part)
        return gx.to(dtype), None, None, None, None, None


class HexPixelShuffle(nn.Module):
    """Sub-pixel up-sampling on the hex lattice ("codes in old versions.txt":68-126): ``(B, C*r*r, H, W)`` ->
    ``(B, C, r*H - r + 1, r*W - r + r//2 ... )`` float32, each group of ``r*r`` channels spread over a radius-``r``
    hexagonal neighbourhood.  The reference needs ``2*(3r^2-3r+1)`` strided assignments into a doubled canvas plus a
    ``torch.cat`` per sub-channel; here the index rule is evaluated once per shape on the host
    (``pixel_shuffle_table``) and the shuffle is one gather launch (its adjoint, one scatter launch, in backward).
    Deviation: ``upscale_factor = 1`` raises (the reference returns an empty tensor from its ``[0:-0]`` crop)."""

    def __init__(self, upscale_factor):
        super().__init__()
        if int(upscale_factor) < 2:
            raise ValueError("upscale_factor must be >= 2")
        self.upscale_factor = int(upscale_factor)

    def forward(self, input: Tensor) -> Tensor:
        input = _as4(input)
        r = self.upscale_factor
        B, Cin, H, W = input.shape
        if Cin % (r * r) != 0:
            raise Exception(f"pixel shuffle needs the channel count ({Cin}) to be a multiple of upscale_factor**2 ({r * r})")
        if input.dtype not in (torch.float32, torch.float64, torch.bfloat16, torch.uint8):
            input = input.float()
        cout = Cin // (r * r)
        table, out_hw = _pixel_shuffle_offsets(r, H, W, cout, str(input.device))
        return _PlaneGatherFn.apply(input, table, out_hw, cout, H * W)

    def extra_repr(self):
        return f"upscale_factor={self.upscale_factor}"


# ------------------------------------------------------------------------------------------------
# hex transposed convolution (retired from the reference: "codes in old versions.txt":129-274)
# ------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=64)
def conv_transpose_tables(radius: int, stride: int, even_odd_offset: int, H: int, W: int):
    """Host-side index rule of ``HexConvTranspose2d`` for an ``H x W`` input.

    The reference (old versions :186-205) paints the input into a zero canvas in doubled (type1) coordinates -- even
    input rows at canvas rows ``2s*a``, sub-columns ``eo*s + 2s*b`` and ``+1``; odd input rows at rows ``s + 2s*a``,
    sub-columns ``(1-eo)*s + 2s*b`` and ``+1`` -- pads it by ``r-1`` rows and ``2(r-1)`` sub-columns, and runs the two
    strided convs of ``HexConv2d`` on it, the odd one starting ``s`` rows lower and ``s`` sub-columns further right
    (:232-243).  That canvas is the type1 raster of a zero-inserted hex lattice ``U`` with some row parity ``o_u``; the
    even output rows are the even rows of ``Y = HexConv2d(stride=1, padding=0)(U)`` and the odd output rows are rows
    ``2(p+k)+1`` (``s = 2k+1``) or ``2(p+k)`` (``s = 2k``) of ``Y`` taken ``k`` cells further right.

    Returns ``(up, (Hu, Wu), o_u, sel, (Ho, Wo), (Hy, Wy))``: ``up`` int64 ``Hu*Wu`` offsets ``row*W + col`` into an input
    plane (-1 = inserted zero), ``sel`` int64 ``Ho*Wo`` offsets into a ``Hy x Wy`` plane of ``Y``.  Raises ValueError
    where the reference's slice assignments do not fit (it raises RuntimeError there)."""
    r, s, eo = int(radius), int(stride), int(even_odd_offset)
    p = r - 1
    w1 = 2 * s * W - s + 2 + (1 - s % 2)                                   # :191
    h1 = s * H - s + 1                                                     # :192
    canvas = np.full((h1, w1), -1, np.int64)
    src = np.arange(H * W, dtype=np.int64).reshape(H, W)
    try:
        canvas[0::2 * s, eo * s:-1:2 * s] = src[0::2]                      # :194-197
        canvas[0::2 * s, eo * s + 1::2 * s] = src[0::2]
        canvas[s::2 * s, (1 - eo) * s:-1:2 * s] = src[1::2]                # :199-202
        canvas[s::2 * s, (1 - eo) * s + 1::2 * s] = src[1::2]
    except ValueError as e:
        raise ValueError(f"HexConvTranspose2d: a {H}x{W} input does not fit the stride-{s} canvas ({e})") from None
    canvas = np.pad(canvas, ((p, p), (2 * p, 2 * p)), constant_values=-1)  # :203-204
    Hu, Wc = canvas.shape
    Wu = (Wc - 1) // 2
    up = o_u = None
    for o in (0, 1):                                                       # which row parity makes the canvas a type1 raster
        shift = (np.arange(Hu) + o) % 2
        cols = shift[:, None] + 2 * np.arange(Wu)[None, :]
        first = np.take_along_axis(canvas, cols, 1)
        again = np.full_like(canvas, -1)
        np.put_along_axis(again, cols, first, 1)
        np.put_along_axis(again, cols + 1, first, 1)
        if np.array_equal(again, canvas):
            up, o_u = first, o
            break
    if up is None:
        raise ValueError("HexConvTranspose2d: the zero-stuffed canvas is not a doubled hex raster")
    k_h, k_w = 2 * r - 1, 4 * r - 3
    if Hu - s < k_h or Wc - 1 - s < k_w:
        raise ValueError(f"HexConvTranspose2d: a {H}x{W} input is too small for radius {r}, stride {s}")
    rows_e, rows_o = (Hu - k_h) // 2 + 1, (Hu - s - k_h) // 2 + 1
    Wo = (Wc - 1 - s - k_w) // 2 + 1
    if rows_e - rows_o not in (0, 1):
        raise ValueError("HexConvTranspose2d: even / odd output rows cannot be interleaved")
    Hy, Wy = _conv_out_shape(Hu, Wu, r, 1, 1, 0)
    k = s // 2
    Ho = rows_e + rows_o
    R = np.arange(Ho)[:, None]
    yrow = np.where(R % 2 == 0, R, 2 * (R // 2 + k) + (1 if s % 2 == 1 else 0))
    ycol = np.arange(Wo)[None, :] + np.where(R % 2 == 0, 0, k)
    if yrow.max() >= Hy or ycol.max() >= Wy:
        raise ValueError("HexConvTranspose2d: selection leaves the stride-1 result")
    sel = (yrow * Wy + ycol).astype(np.int64)
    return up.reshape(-1), (Hu, Wu), o_u, sel.reshape(-1), (Ho, Wo), (Hy, Wy)


@functools.lru_cache(maxsize=64)
def _conv_transpose_device_tables(radius, stride, eo, H, W, device):
    up, hu_wu, o_u, sel, ho_wo, hy_wy = conv_transpose_tables(radius, stride, eo, H, W)
    dev = torch.device(device)
    put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    # for an even stride both output parities read even rows of the stride-1 result: the selection reads some elements
    # twice, so its adjoint is a sum of scatters over injective layers
    return put(up), hu_wu, o_u, put(sel), ho_wo, hy_wy, tuple(put(t) for t in _injective_layers(sel))


class HexConvTranspose2d(nn.Module):
    """Hex transposed convolution as the reference retired it ("codes in old versions.txt":129-274): zero-insert the
    lattice by ``stride``, frame it by ``radius-1`` cells, hex-convolve, interleave.  Same constructor, attribute names
    (``in_channel`` / ``out_channel`` sic) and parameters (``kernel`` ``[out, in/groups, 1, 3r^2-3r+1]``, ``bias``).
    Three launches: one table-driven gather (zero insertion), the ``HexConv2d`` kernels (tcgen05 where eligible), one
    table-driven gather (row / column selection); autograd through all three.  The reference builds its canvas on the CPU
    (:193), so it only ever ran on CPU tensors; here CUDA tensors only.  float32 result (:265-268)."""

    def __init__(self, in_channels, out_channels, even_odd_offset, hexkernel_radius, stride=1, groups=1, bias=False):
        super().__init__()
        self.in_channel = in_channels
        self.out_channel = out_channels
        self.even_odd_offset = even_odd_offset
        self.hexkernel_radius = hexkernel_radius
        self.hexkernel_size = 2 * hexkernel_radius - 1
        self.k_w = 4 * hexkernel_radius - 3
        self.k_h = self.hexkernel_size
        self.kernelnum = 3 * hexkernel_radius ** 2 - 3 * hexkernel_radius + 1
        self.sh = stride
        self.sw = stride * 2
        self.out_even_odd_offset = 0
        self.groups = groups
        self.b = bias
        if in_channels % groups != 0:
            raise ValueError('in_channels must be divisible by groups')
        if out_channels % groups != 0:
            raise ValueError('out_channels must be divisible by groups')
        self.kernel = nn.Parameter(torch.empty([out_channels, in_channels // groups, 1, self.kernelnum], dtype=torch.float))
        if self.b == True:  # noqa: E712
            self.bias = nn.Parameter(torch.empty([out_channels, ]))
        else:
            self.register_parameter('bias', None)
        self.algo = 0
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.kernel, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.kernel)
            if fan_in != 0:
                bound = 1 / math.sqrt(fan_in)
                init.uniform_(self.bias, -bound, bound)

    def forward(self, input: Tensor) -> Tensor:
        input = _as4(input)
        if input.dtype not in (torch.float32, torch.bfloat16, torch.float64):
            input = input.float()
        B, Cin, H, W = input.shape
        up, (Hu, Wu), o_u, sel, (Ho, Wo), (Hy, Wy), sel_layers = _conv_transpose_device_tables(
            self.hexkernel_radius, self.sh, int(self.even_odd_offset) % 2, H, W, str(input.device))
        U = _PlaneGatherFn.apply(input, up, (Hu, Wu), Cin, H * W)                       # zero-inserted, framed lattice
        Y = hexconv2d(U, self.kernel, self.bias, even_odd_offset=o_u, radius=self.hexkernel_radius, stride=1, padding=0,
                      groups=self.groups, algo=self.algo)
        assert tuple(Y.shape[-2:]) == (Hy, Wy)
        return _PlaneGatherFn.apply(Y, sel, (Ho, Wo), self.out_channel, Hy * Wy, sel_layers)

    def __repr__(self):
        return (f"HexConvTranspose2d({self.in_channel}, {self.out_channel}, kernel_radius={self.hexkernel_radius}, "
                f"stride={self.sh}, groups={self.groups}, bias={self.b})")
