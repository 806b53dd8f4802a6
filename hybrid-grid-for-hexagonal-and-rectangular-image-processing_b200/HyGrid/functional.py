"""Batched, device-resident entry points of the rect<->hex path (torch CUDA tensors in/out).

These are the calls the numpy-facing mirrors (``geometry_np`` / ``geometry_torch`` / ``IMAGE`` /
``HEXIMAGE``) are thin shims over.  The reference has no batch dimension; here any number of
leading dimensions is treated as independent planes: ``(..., H, W) -> (..., h1, w1)``.

The sampling geometry of every reference resampler depends only on the shapes, so the 1-D
coordinate tables are produced on the host with *the very same* ``linspace`` call the reference
makes (numpy float64, or float32 ``torch.linspace`` for the geometry_torch twin), cached per
shape and device, and everything after that -- truncation, axial->offset indexing, triangle
selection, weights, gather, blend -- runs in the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import torch

from . import _native as nv

__all__ = ["rect_to_hex", "hex_to_rect", "hex_resize", "hex_warp", "hex_warp_affine",
           "rect2hex_index", "hexsrc_index", "axial_to_offset", "offset_to_axial",
           "hex_to_type1", "hex_to_type2", "type1_to_hex", "type2_to_hex"]

_MATH = {"exact": nv.MATH_EXACT, "fast": nv.MATH_FAST}


# ---------------------------------------------------------------------------------------
# coordinate tables (host, the reference's own linspace calls; cached)
# ---------------------------------------------------------------------------------------
def _linspace(lo, hi, n, twin):
    if twin == "torch":   # geometry_torch.py:252-253: float32 linspace widened to double
        return torch.linspace(lo, hi, n).double().numpy()
    return np.linspace(lo, hi, n)


@functools.lru_cache(maxsize=64)
def _tables(kind, h, w, h1, w1, twin, device):
    if kind == "rect2hex":      # geometry_np.py:401-421
        xs = _linspace(-(h / 2), h / 2, h1, "np")
        ys = _linspace(-(w / 2 + 0.5), w / 2 + 0.5, w1, "np")
    elif kind == "hex2rect":    # geometry_np.py:236-254 / geometry_torch.py:235-253
        xs = _linspace(-(h / 2 - 0.5), h / 2 - 0.5, h1, twin)
        ys = _linspace(-((w + 0.5) / 2 - 0.75), (w + 0.5) / 2 - 0.75, w1, twin)
    elif kind == "hexresize":   # geometry_np.py:560-578
        xs = _linspace(-(h / 2 - 0.5), h / 2 - 0.5, h1, "np")
        ys = _linspace(-((w + 0.5) / 2 - 0.5), (w + 0.5) / 2 - 0.5, w1, "np")
    else:
        raise KeyError(kind)
    dev = torch.device(device)
    xs, ys = np.ascontiguousarray(xs, dtype=np.float64), np.ascontiguousarray(ys, dtype=np.float64)
    return torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev), xs, ys


def coordinate_tables(kind, h, w, h1, w1, twin="np", device="cuda"):
    """(xs_dev, ys_dev, xs_host, ys_host) float64 tables of one resampler."""
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return _tables(kind, int(h), int(w), int(h1), int(w1), twin, str(dev))


def _planes(x):
    if x.dim() < 2:
        raise Exception(f"dim of image should be 2 or more, but got dim = {x.dim()} instead")
    nv.require_cuda(x, "image")
    x = x.contiguous()
    h, w = x.shape[-2:]
    planes = x.numel() // (h * w) if h * w else 0
    return x, planes, h, w


def _result(out, shape, dtype, like):
    """The result buffer: a fresh one, or the caller's ``out`` after checking that the kernel may write it."""
    if out is None:
        return torch.empty(shape, dtype=dtype, device=like.device)
    if tuple(out.shape) != tuple(shape) or out.device != like.device or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous {tuple(shape)} tensor on {like.device}, "
                         f"got {tuple(out.shape)} on {out.device} (contiguous={out.is_contiguous()})")
    if dtype is not None and out.dtype != dtype:
        raise TypeError(f"out must have dtype {dtype}, got {out.dtype}")
    return out


def _out_dtype(x, out_dtype):
    if out_dtype is None:   # numpy promotion of float64 weights with the image dtype
        return torch.float64
    return out_dtype


# ---------------------------------------------------------------------------------------
# R1 rect -> hex
# ---------------------------------------------------------------------------------------
def rect_to_hex(x: torch.Tensor, hex_dsize=None, interpolation="nearest", out_dtype=None, math="exact",
                out: torch.Tensor | None = None) -> torch.Tensor:
    """``geometry_np.rect_to_hex_resample`` (geometry_np.py:358-519) for ``(..., H, W)`` CUDA tensors.

    nearest keeps the element type.  bilinear returns ``out_dtype`` (default float64, the
    reference's result type; pass ``torch.float32`` to halve the write traffic).
    ``math='exact'``: float64 arithmetic in the reference's operation order (bit-identical for a
    float64 result); ``'fast'``: float32 FMAs."""
    method = {"nearest": 0, "bilinear": 1}[interpolation]
    x, planes, h, w = _planes(x)
    h1, w1 = (h, w) if hex_dsize is None else (int(hex_dsize[0]), int(hex_dsize[1]))
    xs, ys, hxs, hys = coordinate_tables("rect2hex", h, w, h1, w1, "np", x.device)
    shape = x.shape[:-2] + (h1, w1)
    st = nv.stream_ptr(x.device)
    if method == 0:
        y = _result(out, shape, x.dtype, x)
        nv.call("hg_rect2hex_nearest", nv.ptr(x), nv.ptr(y), nv.ptr(xs), nv.ptr(ys), planes, h, w, h1, w1,
                x.element_size(), st)
    else:
        y = _result(out, shape, _out_dtype(x, out_dtype) if out is None or out_dtype is not None else None, x)
        nv.call("hg_rect2hex_bilinear", nv.ptr(x), nv.ptr(y), nv.ptr(xs), nv.ptr(ys), C.c_void_p(hxs.ctypes.data),
                C.c_void_p(hys.ctypes.data), planes, h, w, h1, w1,
                nv.hg_dtype(x.dtype), nv.hg_dtype(y.dtype), _MATH[math], st)
    return y


def rect2hex_index(h, w, h1, w1, device="cuda"):
    """Integer/fraction tables (i_n, i_f, j_n, j_f) of rect->hex (geometry_np.py:440-449)."""
    xs, ys, _, _ = coordinate_tables("rect2hex", h, w, h1, w1, "np", device)
    dev = xs.device
    i_n = torch.empty(h1, dtype=torch.int32, device=dev); i_f = torch.empty(h1, dtype=torch.float64, device=dev)
    j_n = torch.empty(w1, dtype=torch.int32, device=dev); j_f = torch.empty(w1, dtype=torch.float64, device=dev)
    nv.call("hg_rect2hex_index", nv.ptr(xs), nv.ptr(ys), h, w, h1, w1, nv.ptr(i_n), nv.ptr(i_f), nv.ptr(j_n),
            nv.ptr(j_f), nv.stream_ptr(dev))
    return i_n, i_f, j_n, j_f


# ---------------------------------------------------------------------------------------
# R2 / R4 hex -> rect, hex -> hex resize
# ---------------------------------------------------------------------------------------
def _hexsrc(kind, x, dsize, interpolation, out_dtype, math, twin, out):
    method = {"nearest": 0, "linear": 1, "bilinear": 2}[interpolation]
    if method == 2:
        raise NotImplementedError("'bilinear' on a hex-lattice source executes no branch in the reference "
                                  "(geometry_np.py:333-356 returns uninitialised memory)")
    x, planes, h, w = _planes(x)
    h1, w1 = (h, w) if dsize is None else (int(dsize[0]), int(dsize[1]))
    xs, ys, hxs, hys = coordinate_tables(kind, h, w, h1, w1, twin, x.device)
    shape = x.shape[:-2] + (h1, w1)
    st = nv.stream_ptr(x.device)
    if method == 0:
        y = _result(out, shape, x.dtype, x)
        nv.call("hg_hex2rect_nearest", nv.ptr(x), nv.ptr(y), nv.ptr(xs), nv.ptr(ys), planes, h, w, h1, w1,
                x.element_size(), st)
    else:
        y = _result(out, shape, _out_dtype(x, out_dtype) if out is None or out_dtype is not None else None, x)
        nv.call("hg_hex2rect_linear", nv.ptr(x), nv.ptr(y), nv.ptr(xs), nv.ptr(ys), C.c_void_p(hxs.ctypes.data),
                C.c_void_p(hys.ctypes.data), planes, h, w, h1, w1,
                nv.hg_dtype(x.dtype), nv.hg_dtype(y.dtype), _MATH[math], st)
    return y


def hex_to_rect(x, rect_dsize=None, interpolation="nearest", out_dtype=None, math="exact", twin="torch", out=None):
    """``geometry_torch.hex_to_square_resample`` (geometry_torch.py:191-358; ``twin='torch'``: float32
    linspace) / ``geometry_np.hex_to_rect_resample`` (geometry_np.py:191-356; ``twin='np'``)."""
    return _hexsrc("hex2rect", x, rect_dsize, interpolation, out_dtype, math, twin, out)


def hex_resize(x, dsize, interpolation="linear", out_dtype=None, math="exact", out=None):
    """``geometry_np.hexresize`` (geometry_np.py:520-681)."""
    return _hexsrc("hexresize", x, dsize, interpolation, out_dtype, math, "np", out)


def hexsrc_index(h, w, xs, ys, coords_2d=False, coord_f32=False):
    """Per-sample integer tables of the hex-source resamplers: (i_n, j_n, tri, off[3])."""
    dev = xs.device
    if coords_2d:
        h1, w1 = xs.shape
    else:
        h1, w1 = xs.numel(), ys.numel()
    i_n = torch.empty((h1, w1), dtype=torch.int32, device=dev)
    j_n = torch.empty((h1, w1), dtype=torch.int32, device=dev)
    tri = torch.empty((h1, w1), dtype=torch.uint8, device=dev)
    off = torch.empty((3, h1, w1), dtype=torch.int32, device=dev)
    nv.call("hg_hexsrc_index", nv.ptr(xs.contiguous()), nv.ptr(ys.contiguous()), int(coords_2d), int(coord_f32),
            h, w, h1, w1, nv.ptr(i_n), nv.ptr(j_n), nv.ptr(tri), nv.ptr(off), nv.stream_ptr(dev))
    return i_n, j_n, tri, off


# ---------------------------------------------------------------------------------------
# R3 hex -> hex affine warp
# ---------------------------------------------------------------------------------------
def warp_lattice(h, w, H, twin="torch"):
    """Output lattice of the warp (rows, cols, inv(H)) exactly as the reference builds it
    (geometry_torch.py:56-98 / geometry_np.py:56-100)."""
    H = np.asarray(H, dtype=np.float64)
    cx, cy = h / 2 - 0.5, (w + 0.5) / 2 - 0.5
    corners = [[-cx, -cy, 1.0], [-cx, cy, 1.0], [cx, -cy, 1.0], [cx, cy, 1.0]]
    if twin == "torch":
        Ht = torch.tensor(H).to(torch.float64)
        tc = torch.matmul(Ht, torch.tensor(corners, dtype=torch.double).T)
        lo0, lo1 = torch.min(tc[0]).item(), torch.min(tc[1]).item()
        hi0, hi1 = torch.max(tc[0]).item(), torch.max(tc[1]).item()
        rows = torch.arange(lo0, hi0 + 1, 1).double().numpy()
        cols = torch.arange(lo1, hi1 + 0.5, 1).double().numpy()
        Hi = torch.linalg.inv(Ht).numpy()
    elif twin == "numba":
        # legacy HyGrid/geometry.py:208-221: integer-truncated start, and the ROW extent reused for the columns
        tc = np.matmul(H, np.array(corners).T)
        lo0, hi0 = tc[0].min(), tc[0].max()
        rows = np.arange(int(lo0), hi0 + 1, 1.0)
        cols = np.arange(int(lo0), hi0 + 0.5, 1.0)
        Hi = np.linalg.inv(H)
    else:
        tc = np.matmul(H, np.array(corners).T)
        rows = np.arange(tc[0].min(), tc[0].max() + 1, 1)
        cols = np.arange(tc[1].min(), tc[1].max() + 0.5, 1)
        Hi = np.linalg.inv(H)
    return rows, cols, Hi


def _warp_planes(rows, cols, Hi, twin):
    """Inverse-mapped coordinate planes with the reference's own contraction call, so that the
    coordinates (and through them every integer index) are bit-identical."""
    h1, w1 = rows.shape[0], cols.shape[0]
    X = np.broadcast_to(rows[:, None], (h1, w1)).copy()
    Y = np.broadcast_to(cols[None, :], (h1, w1)).copy()
    Y[1::2] += 0.5
    hom = np.stack([X, Y, np.ones_like(X)], 0)
    if twin == "torch":
        inv = torch.einsum("ij, jkl -> ikl", torch.tensor(Hi), torch.tensor(hom)).to(torch.float)
        return inv[0].contiguous(), inv[1].contiguous()
    inv = np.einsum("ij, jkl -> ikl", Hi, hom)
    return torch.from_numpy(np.ascontiguousarray(inv[0])), torch.from_numpy(np.ascontiguousarray(inv[1]))


def hex_warp(x, H=np.eye(3), interpolation="nearest", out_dtype=None, twin="torch"):
    """``geometry_torch.image_geometric_transformation_gpu`` (geometry_torch.py:7-189, ``twin='torch'``,
    float32 coordinates) / ``geometry_np.image_geometric_transformation`` (geometry_np.py:6-189, ``twin='np'``) /
    the legacy numba ``geometry.image_geometric_transformation_gpu`` (geometry.py:156-262, ``twin='numba'``).
    Coordinates are inverse-mapped on the host like the reference does; the rest runs on the GPU."""
    method = {"nearest": 0, "linear": 1, "bilinear": 2}[interpolation]
    if method == 2:
        raise NotImplementedError("'bilinear' executes no branch in the reference")
    x, planes, h, w = _planes(x)
    rows, cols, Hi = warp_lattice(h, w, H, twin)
    cx, cy = _warp_planes(rows, cols, Hi, twin)
    cx, cy = cx.to(x.device), cy.to(x.device)
    h1, w1 = cx.shape
    f32 = int(twin == "torch")
    shape = x.shape[:-2] + (h1, w1)
    st = nv.stream_ptr(x.device)
    if method == 0:
        y = torch.empty(shape, dtype=x.dtype, device=x.device)
        nv.call("hg_hexwarp_nearest", nv.ptr(x), nv.ptr(y), nv.ptr(cx), nv.ptr(cy), f32, planes, h, w, h1, w1,
                x.element_size(), st)
    else:
        if out_dtype is None:   # torch promotion: float32 weights stay float32 unless the image is float64
            out_dtype = torch.float32 if (f32 and x.dtype != torch.float64) else torch.float64
        y = torch.empty(shape, dtype=out_dtype, device=x.device)
        nv.call("hg_hexwarp_linear", nv.ptr(x), nv.ptr(y), nv.ptr(cx), nv.ptr(cy), f32, planes, h, w, h1, w1,
                nv.hg_dtype(x.dtype), nv.hg_dtype(y.dtype), st)
    return y


def hex_warp_affine(x, H=np.eye(3), interpolation="linear", out_dtype=None, coord_f32=True):
    """Same warp with the inverse map evaluated inside the kernel (no coordinate planes in HBM, no
    host einsum): the high-throughput path for batches.  Coordinates can differ from the host
    einsum in the last ulp, so this path is toleranced, not bit-pinned."""
    method = {"nearest": 0, "linear": 1}[interpolation]
    x, planes, h, w = _planes(x)
    rows, cols, Hi = warp_lattice(h, w, H, "torch" if coord_f32 else "np")
    h1, w1 = rows.shape[0], cols.shape[0]
    hinv = (C.c_double * 6)(*[float(v) for v in np.asarray(Hi)[:2].reshape(-1)])
    if method == 0:
        out_dtype = x.dtype
    elif out_dtype is None:
        out_dtype = torch.float32 if (coord_f32 and x.dtype != torch.float64) else torch.float64
    y = torch.empty(x.shape[:-2] + (h1, w1), dtype=out_dtype, device=x.device)
    nv.call("hg_hexwarp_affine", nv.ptr(x), nv.ptr(y), hinv, float(rows[0]) if h1 else 0.0,
            float(cols[0]) if w1 else 0.0, int(coord_f32), method, planes, h, w, h1, w1,
            nv.hg_dtype(x.dtype), nv.hg_dtype(y.dtype), nv.stream_ptr(x.device))
    return y


# ---------------------------------------------------------------------------------------
# lattice index helpers
# ---------------------------------------------------------------------------------------
def axial_to_offset(i: torch.Tensor, j_ax: torch.Tensor) -> torch.Tensor:
    """j_off = j_ax - trunc((i+1)/2)  (geometry_np.py:288-295)."""
    i = nv.require_cuda(i).to(torch.int32).contiguous(); j = j_ax.to(torch.int32).contiguous()
    i, j = torch.broadcast_tensors(i, j)
    i, j = i.contiguous(), j.contiguous()
    out = torch.empty_like(j)
    nv.call("hg_axial_to_offset_i32", nv.ptr(i), nv.ptr(j), nv.ptr(out), j.numel(), nv.stream_ptr(j.device))
    return out


def offset_to_axial(i: torch.Tensor, j_off: torch.Tensor) -> torch.Tensor:
    i = nv.require_cuda(i).to(torch.int32).contiguous(); j = j_off.to(torch.int32).contiguous()
    i, j = torch.broadcast_tensors(i, j)
    i, j = i.contiguous(), j.contiguous()
    out = torch.empty_like(j)
    nv.call("hg_offset_to_axial_i32", nv.ptr(i), nv.ptr(j), nv.ptr(out), j.numel(), nv.stream_ptr(j.device))
    return out


# ---------------------------------------------------------------------------------------
# R5 doubled rasters
# ---------------------------------------------------------------------------------------
def _to_type(name, x, even_odd_offset, out_dtype, rows_mul):
    x, planes, H, W = _planes(x)
    out_dtype = out_dtype or x.dtype
    y = torch.empty(x.shape[:-2] + (rows_mul * H, 2 * W + 1), dtype=out_dtype, device=x.device)
    nv.call(name, nv.ptr(x), nv.ptr(y), planes, H, W, int(even_odd_offset) % 2, nv.hg_dtype(x.dtype),
            nv.hg_dtype(out_dtype), nv.stream_ptr(x.device))
    return y


def hex_to_type1(x, even_odd_offset=0, out_dtype=None):
    """(..., H, W) -> (..., H, 2W+1) doubled raster (HexImage.py:139-153 / HexFrames.py:417-445)."""
    return _to_type("hg_hex_to_type1", x, even_odd_offset, out_dtype, 1)


def hex_to_type2(x, even_odd_offset=0, out_dtype=None):
    """(..., H, W) -> (..., 2H, 2W+1) (HexImage.py:154-170 / HexFrames.py:446-449)."""
    return _to_type("hg_hex_to_type2", x, even_odd_offset, out_dtype, 2)


def _from_type(x, rows_step, out_dtype):
    x, planes, Ht, Wt = _planes(x)
    out_dtype = out_dtype or x.dtype
    y = torch.empty(x.shape[:-2] + ((Ht + rows_step - 1) // rows_step, (Wt - 1) // 2), dtype=out_dtype, device=x.device)
    nv.call("hg_type_to_hex", nv.ptr(x), nv.ptr(y), planes, Ht, Wt, rows_step, nv.hg_dtype(x.dtype),
            nv.hg_dtype(out_dtype), nv.stream_ptr(x.device))
    return y


def type1_to_hex(x, out_dtype=None):
    """``[..., 1:-1:2]`` (HexImage.py:109)."""
    return _from_type(x, 1, out_dtype)


def type2_to_hex(x, out_dtype=None):
    """``[..., ::2, 1:-1:2]`` (HexImage.py:111)."""
    return _from_type(x, 2, out_dtype)
