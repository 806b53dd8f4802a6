"""Learned lattice resamplers -- the three layers the reference retired into ``HyGrid/codes in old versions.txt``
(SURVEY.md section 8f rank 3), on one sm_100a kernel family (``hg_dwtaps_*``: depthwise weighted tap gathers on doubled
coordinates, csrc/hg_dwtaps.cu):

* ``Hex_to_Square_Conv2d_by_Double_Stride`` (:1-66)     hex lattice -> square grid, down-sampling by an even factor f
* ``Square_to_Hex_Conv2d_by_Double_Stride`` (:421-493)  square grid -> hex lattice (the reference only runs for f = 2)
* ``Hex_to_Square_original_resolution``     (:587-636)  hex lattice -> square grid at the same resolution

Same constructors, attribute names, parameter name (``kernel``) / shape and initial weights as the retired classes.  The
reference materialises the doubled image, unfolds every window with one strided slice + ``torch.cat`` per tap and loops
over the channels in Python; here each forward is one launch (two for the original-resolution layer) and the backward
two or three.  float32 CUDA tensors; constant padding is virtual, other padding modes go through the pad kernel."""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import _native as nv
from .HexFrames import _as4, pad

__all__ = ["Hex_to_Square_Conv2d_by_Double_Stride", "Square_to_Hex_Conv2d_by_Double_Stride",
           "Hex_to_Square_original_resolution", "resampler_weight"]


def resampler_weight(f: int, kind: str) -> torch.Tensor:
    """Initial f x f weights: inverse distance of every window cell to the output sample, normalised
    (old versions :36-49 'hex_to_square', :445-459 'square_to_hex', :616-623 'original_resolution')."""
    x = torch.arange(0, f).float()
    coor = torch.cartesian_prod(x, x).view(f, f, 2)
    a, b = coor[:, :, 0], coor[:, :, 1]
    if kind == "hex_to_square":
        d2 = (a - (f - 1) / 2) * (a - (f - 1) / 2) + (0.5 * a + b - 3 * (f - 1) / 4) * (0.5 * a + b - 3 * (f - 1) / 4)
    elif kind == "square_to_hex":
        d2 = (a - (f - 1) / 2) * (a - (f - 1) / 2) + (b - (f - 1) / 2) * (b - (f - 1) / 2)
    elif kind == "original_resolution":
        d2 = (a + b - (f - 1)) * (a + b - (f - 1)) + (0.5 * a - 0.5 * b) * (0.5 * a - 0.5 * b)
    else:
        raise KeyError(kind)
    dist = 1 / torch.sqrt(d2)
    return dist / dist.sum()


def _taps(T_rows_cols, sx, doubled):
    ts = nv.TapsSet()
    ts.T, ts.sx, ts.doubled = len(T_rows_cols), sx, int(doubled)
    for t, (r, c) in enumerate(T_rows_cols):
        ts.ry[t], ts.ex[t] = r, c
    return ts


def _desc(x, Ho, Wo, sy, odd_limit, parity, pad_, pad_value, even, odd):
    N, Cc, H, W = x.shape
    d = nv.DwTapsDesc()
    d.N, d.C, d.H, d.W, d.Ho, d.Wo = N, Cc, H, W, Ho, Wo
    d.sy, d.odd_limit, d.parity, d.pad, d.pad_value = sy, odd_limit, parity, pad_, float(pad_value)
    d.set[0], d.set[1] = even, odd
    return d


class _TapGatherFn(torch.autograd.Function):
    """``y = hg_dwtaps_fwd(x, w_even, w_odd)``; ``geom = (Ho, Wo, sy, odd_limit, parity, pad, pad_value, even, odd, shared)``:
    ``shared`` -- both tap sets use the one weight tensor ``w_even`` (its gradient is the sum over both sets)."""

    @staticmethod
    def forward(ctx, x, w_even, w_odd, geom):
        Ho, Wo, sy, odd_limit, parity, pad_, pad_value, even, odd, shared = geom
        x = nv.require_cuda(x, "input").float().contiguous()
        we = w_even.detach().float().contiguous() if w_even is not None else None
        wo = we if shared else (w_odd.detach().float().contiguous() if w_odd is not None else None)
        d = _desc(x, Ho, Wo, sy, odd_limit, parity, pad_, pad_value, even, odd)
        y = torch.empty((x.shape[0], x.shape[1], Ho, Wo), dtype=torch.float32, device=x.device)
        nv.call("hg_dwtaps_fwd", C.byref(d), nv.ptr(x), nv.ptr(we), nv.ptr(wo), nv.ptr(y), nv.stream_ptr(x.device))
        ctx.save_for_backward(x, we, wo)
        ctx.geom = geom
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, we, wo = ctx.saved_tensors
        Ho, Wo, sy, odd_limit, parity, pad_, pad_value, even, odd, shared = ctx.geom
        gy = gy.float().contiguous()
        d = _desc(x, Ho, Wo, sy, odd_limit, parity, pad_, pad_value, even, odd)
        st = nv.stream_ptr(x.device)
        gx = gwe = gwo = None
        if ctx.needs_input_grad[0]:
            gx = torch.zeros_like(x)
            nv.call("hg_dwtaps_dgrad", C.byref(d), nv.ptr(gy), nv.ptr(we), nv.ptr(wo), nv.ptr(gx), st)
        if we is not None and ctx.needs_input_grad[1]:
            gwe = torch.zeros_like(we)
            nv.call("hg_dwtaps_wgrad", C.byref(d), nv.ptr(x), nv.ptr(gy), nv.ptr(gwe), 0, st)
            if shared:
                nv.call("hg_dwtaps_wgrad", C.byref(d), nv.ptr(x), nv.ptr(gy), nv.ptr(gwe), 1, st)
        if not shared and wo is not None and ctx.needs_input_grad[2]:
            gwo = torch.zeros_like(wo)
            nv.call("hg_dwtaps_wgrad", C.byref(d), nv.ptr(x), nv.ptr(gy), nv.ptr(gwo), 1, st)
        return gx, gwe, gwo, None


def _padded(x, padding, mode, value):
    """(input, virtual pad, pad value): constant padding stays virtual, other modes run the pad kernel."""
    x = _as4(x)
    if padding and mode != 'constant':
        return pad(x, padding, mode, value), 0, 0.0
    return x, int(padding), float(value or 0)


class Hex_to_Square_Conv2d_by_Double_Stride(nn.Module):
    """Hex lattice -> square grid, down-sampled by an even ``downsample_factor`` f (old versions :1-66): a depthwise f x f
    rhombus window -- tap (i, m) at doubled column i + 2m of window row i -- moved f rows and 2f - 1 doubled columns per
    output.  ``kernel`` (C, f, f), initialised to normalised inverse distances."""

    def __init__(self, channels, even_odd_offset, downsample_factor=2, padding=0, padding_mode='constant', padding_value=0):
        super().__init__()
        self.in_channels = channels
        self.out_channels = channels
        if downsample_factor % 2 != 0:
            raise Exception("降采样因子必须是2的倍数")
        self.even_odd_offset = even_odd_offset
        self.padded_even_odd_offset = (even_odd_offset + padding) % 2
        self.kernel_size = downsample_factor
        self.kernelnum = downsample_factor ** 2
        self.out_even_odd_offset = 0
        self.pad = padding
        self.padding_mode = padding_mode
        self.padding_value = padding_value
        self.kernel = nn.Parameter(self.generate_weight(downsample_factor))
        self.k_h = self.kernel_size
        self.k_w = 3 * downsample_factor - 2
        self.stride = (downsample_factor, 2 * downsample_factor - 1)
        self.downsample_factor = downsample_factor

    def generate_weight(self, f):
        return resampler_weight(f, "hex_to_square").unsqueeze(0).repeat(self.out_channels, 1, 1)

    def forward(self, input):
        f = self.downsample_factor
        x, p, pv = _padded(input, self.pad, self.padding_mode, self.padding_value)
        Hp, Wp = x.shape[2] + 2 * p, x.shape[3] + 2 * p
        o = self.padded_even_odd_offset
        wt = 2 * Wp if o % 2 == 0 else 2 * Wp - 1                       # type1[..., 1:] or [..., 1:-1]  (:60)
        Ho, Wo = (Hp - f) // f + 1, (wt - self.k_w) // (2 * f - 1) + 1
        if Hp < f or wt < self.k_w:
            raise RuntimeError(f"input {tuple(x.shape[2:])} is smaller than the {f} x {self.k_w} window")
        taps = _taps([(i, 1 + i + 2 * m) for i in range(f) for m in range(f)], 2 * f - 1, True)
        geom = (Ho, Wo, f, 0, o, p, pv, taps, taps, True)
        return _TapGatherFn.apply(x, self.kernel.reshape(self.in_channels, f * f), None, geom)

    def __repr__(self):
        return (f"Hex_to_Square_Conv2d_by_Double_Stride({self.in_channels}, downsample_factor=stride*2={self.downsample_factor}, "
                f"padding={self.pad})")


class Square_to_Hex_Conv2d_by_Double_Stride(nn.Module):
    """Square grid -> hex lattice (old versions :421-493): a learned 2 x 2 box, rows 2R and 2R + 1, whose odd output rows
    start one pixel (half a hex cell) further right.  ``kernel`` (C, f*f).  The reference's unfold is hard-wired to a
    2 x 2 window (:468-469), so only ``downsample_factor = 2`` runs there; the same restriction applies here."""

    def __init__(self, channels, downsample_factor, padding=0, padding_mode='constant', padding_value=0):
        super().__init__()
        if downsample_factor % 2 != 0:
            raise Exception("降采样因子必须是2的倍数")
        self.in_channels = channels
        self.out_channels = channels
        self.kernel_size = downsample_factor
        self.stride = (downsample_factor, downsample_factor)
        self.out_even_odd_offset = 0
        self.pad = padding
        self.padding_mode = padding_mode
        self.padding_value = padding_value
        self.kernel = nn.Parameter(self.generate_weight(downsample_factor))
        self.downsample_factor = downsample_factor

    def generate_weight(self, f):
        return resampler_weight(f, "square_to_hex").unsqueeze(0).repeat(self.out_channels, 1, 1).view(self.in_channels, f * f)

    def forward(self, input):
        if self.downsample_factor != 2:
            raise RuntimeError("size mismatch: the 2 x 2 unfold of the reference only matches downsample_factor = 2 "
                               "(codes in old versions.txt:468-476)")
        x, p, pv = _padded(input, self.pad, self.padding_mode, self.padding_value)
        Hp, Wp = x.shape[2] + 2 * p, x.shape[3] + 2 * p
        he, ho = math.ceil((Hp - 1) / 4), math.ceil((Hp - 3) / 4)
        Wo = int((Wp - 2) / 2)
        if he - ho not in (0, 1) or he < 1 or Wo < 1:
            raise RuntimeError("even / odd rows cannot be interleaved (shape mismatch)")
        box = [(i, j) for i in range(2) for j in range(2)]
        even, odd = _taps(box, 2, False), _taps([(i, j + 1) for i, j in box], 2, False)
        geom = (he + ho, Wo, 2, 1 << 30, 0, p, pv, even, odd, True)
        return _TapGatherFn.apply(x, self.kernel, None, geom)

    def __repr__(self):
        return (f"Square_to_Hex_Conv2d_by_Double_Stride({self.in_channels}, {self.out_channels}, kernel_radius={self.kernel_size}, "
                f"downsample_factor={self.downsample_factor}, padding={self.pad})")


class Hex_to_Square_original_resolution(nn.Module):
    """Hex lattice -> square grid at the same resolution (old versions :587-636): even rows are kept; odd rows 1, 3, ...
    (< H - 1), which sit half a cell to the side, are re-interpolated at the even rows' column positions from the rhombus
    {(R-1, 2J+2), (R, 2J+1), (R, 2J+3), (R+1, 2J+2)} of the doubled view; the first column is dropped.  ``kernel`` (C, 4),
    frozen unless ``trainable``."""

    def __init__(self, channels, even_odd_offset, padding=0, padding_mode='constant', padding_value=0, trainable=False):
        super().__init__()
        self.in_channels = channels
        self.out_channels = channels
        self.even_odd_offset = even_odd_offset
        self.offset = (even_odd_offset + padding) % 2
        self.kernel_size = 2
        self.kernelnum = 4
        self.out_even_odd_offset = 0
        self.padding = padding
        self.padding_mode = padding_mode
        self.padding_value = padding_value
        self.kernel = nn.Parameter(self.generate_weight(2), requires_grad=trainable)

    def generate_weight(self, f):
        return resampler_weight(f, "original_resolution").unsqueeze(0).repeat(self.out_channels, 1, 1).view(self.in_channels, f * f)

    def forward(self, input):
        x, p, pv = _padded(input, self.padding, self.padding_mode, self.padding_value)
        Hp, Wp = x.shape[2] + 2 * p, x.shape[3] + 2 * p
        if Hp < 3:
            raise RuntimeError("fewer than three rows: the window of the odd rows does not fit")
        keep = _taps([(0, 1)], 1, False)                                       # even rows: the padded image without column 0
        rhombus = _taps([(-1, 2), (0, 1), (0, 3), (1, 2)], 2, True)
        geom = (Hp, Wp - 1, 1, Hp - 1, self.offset, p, pv, keep, rhombus, False)
        return _TapGatherFn.apply(x, None, self.kernel, geom)
