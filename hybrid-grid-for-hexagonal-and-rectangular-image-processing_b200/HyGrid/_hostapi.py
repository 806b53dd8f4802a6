"""numpy-in / numpy-out plumbing shared by ``geometry_np`` and ``geometry_torch``: host arrays go through
the C ABI host entry points (pinned, chunked H2D / kernel / D2H pipeline in libhygrid_b200.so)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nv
from . import functional as Fn

_INTERP_SRC = (np.dtype(np.uint8), np.dtype(np.float32), np.dtype(np.float64))


def as_chw(img):
    """(C,H,W) view of a 2-D / 3-D array; the reference's dimension error otherwise
    (geometry_np.py:19-25, :199-205, :365-371)."""
    img = np.asarray(img)
    if img.ndim == 3:
        return img
    if img.ndim == 2:
        return img[None]
    raise Exception(f"dim of image should be 2 or 3, but got dim = {img.ndim} instead")


def device_index(device=None) -> int:
    if not torch.cuda.is_available():
        raise nv.HyGridNativeError("no CUDA device: HyGrid on B200 has no CPU path")
    if device is None:
        return torch.cuda.current_device()
    d = torch.device(device)
    return d.index if d.index is not None else torch.cuda.current_device()


def resample_host(kind, img, dsize, interp, out_dtype, math, twin, device=None):
    """kind 'rect2hex' | 'hex2rect' | 'hexresize'; interp 0 nearest, 1 interpolating.  Returns (C,h1,w1)."""
    img = as_chw(img)
    c, h, w = img.shape
    h1, w1 = (h, w) if dsize is None else (int(dsize[0]), int(dsize[1]))
    dev = device_index(device)
    _, _, xs, ys = Fn.coordinate_tables(kind, h, w, h1, w1, twin, torch.device("cuda", dev))
    xs, ys = np.ascontiguousarray(xs), np.ascontiguousarray(ys)
    if interp == 0:
        src = np.ascontiguousarray(img)
        if src.dtype.itemsize not in (1, 2, 4, 8):
            raise TypeError(f"unsupported array dtype {src.dtype}")
        dst = np.empty((c, h1, w1), dtype=src.dtype)
        sdt = ddt = {1: nv.U8, 2: nv.U16, 4: nv.F32, 8: nv.F64}[src.dtype.itemsize]
    else:
        src = img if img.dtype in _INTERP_SRC else img.astype(np.float64)   # integers widen exactly
        src = np.ascontiguousarray(src)
        dst = np.empty((c, h1, w1), dtype=out_dtype)
        sdt, ddt = nv.hg_dtype(src.dtype), nv.hg_dtype(dst.dtype)
    name = "hg_host_rect2hex" if kind == "rect2hex" else "hg_host_hex2rect"
    nv.call(name, C.c_void_p(src.ctypes.data), C.c_void_p(dst.ctypes.data), C.c_void_p(xs.ctypes.data),
            C.c_void_p(ys.ctypes.data), c, h, w, h1, w1, sdt, ddt, interp, math, dev)
    return dst


def to_device(img, device=None):
    img = as_chw(img)
    if img.dtype not in _INTERP_SRC:
        img = img.astype(np.float64)
    return torch.from_numpy(np.ascontiguousarray(img)).to(torch.device("cuda", device_index(device)))
