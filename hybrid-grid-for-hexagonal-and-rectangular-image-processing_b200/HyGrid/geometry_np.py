"""``HyGrid.geometry_np`` on B200: the reference's numpy-in / numpy-out resamplers with the same names,
arguments, interpolation names, result dtypes and ``.squeeze()``d ``(C, h1, w1)`` results
(/root/reference/HyGrid/geometry_np.py), computed by the sm_100a kernels of libhygrid_b200.so.

The coordinate tables come from the very same ``np.linspace`` calls as the reference and the kernels
evaluate the rest in float64 in the reference's operation order, so interpolated float64 results are
bit-identical to the reference, and the integer lattice indexing is exact.

Deviations: ``'nearest'`` on a hex source works (the reference raises from a
``np.min`` unpacking bug, geometry_np.py:172/339/664) and follows the working geometry_torch rule;
``'bilinear'`` on a hex source raises NotImplementedError (the reference returns uninitialised memory);
2-D inputs are accepted as one band; ``heximpad`` works (the reference forgets ``import numbers``) and keeps the
``cv2.copyMakeBorder`` semantics of the reference's call, quirks included (a scalar ``pad_val`` reaches band 0 only).
The ``offset`` argument is dead in the reference too (it only perturbs a variable that is never read).
"""
from __future__ import annotations

import numbers
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from . import _native as nv
from . import functional as Fn
from ._hostapi import as_chw, resample_host, to_device

__all__ = ["image_geometric_transformation", "hex_to_rect_resample", "rect_to_hex_resample", "hexresize",
           "heximpad", "hex_impad_to_multiple"]

_HEX_METHODS = {'nearest': 0, 'linear': 1, 'bilinear': 2}


def _hex_method(interpolation):
    method = _HEX_METHODS[interpolation]            # KeyError like the reference's method_dict lookup
    if method == 2:
        raise NotImplementedError("'bilinear' on a hex-lattice source executes no branch in the reference "
                                  "(geometry_np.py:333-356)")
    return method


def image_geometric_transformation(img: np.array, H: np.array = np.eye(3), interpolation='nearest', offset=0) -> np.array:
    """hex -> hex affine warp (geometry_np.py:6-189): float64 coordinates."""
    _hex_method(interpolation)
    x = to_device(img)
    out = Fn.hex_warp(x, H, interpolation, twin="np")
    res = out.cpu().numpy()
    if interpolation == 'nearest':
        res = res.astype(as_chw(img).dtype, copy=False)
    return res.squeeze()


def hex_to_rect_resample(hex_image, rect_dsize=None, interpolation='nearest', offset=0):
    """geometry_np.py:191-356."""
    method = _hex_method(interpolation)
    return resample_host("hex2rect", hex_image, rect_dsize, method, np.float64, nv.MATH_EXACT, "np").squeeze()


def rect_to_hex_resample(rect_image, hex_dsize=None, interpolation='nearest', offset=0):
    """geometry_np.py:358-519 (interpolation names: 'nearest', 'bilinear')."""
    method = {'nearest': 0, 'bilinear': 1}[interpolation]
    return resample_host("rect2hex", rect_image, hex_dsize, method, np.float64, nv.MATH_EXACT, "np").squeeze()


def hexresize(image, dsize, interpolation="linear", offset=0):
    """geometry_np.py:520-681."""
    method = _hex_method(interpolation)
    return resample_host("hexresize", image, dsize, method, np.float64, nv.MATH_EXACT, "np").squeeze()


def heximpad(img: np.ndarray, *, shape: Optional[Tuple[int, int]] = None, padding: Union[int, tuple, None] = None,
             pad_val: Union[float, List] = 0, padding_mode: str = 'constant') -> np.ndarray:
    """geometry_np.py:683-732: pad an (H, W[, C]) image; the top pad is rounded down to an even number of
    rows (and the remainder moved to the bottom) so that the row parity of the hex lattice is kept.
    Limitation: ``'reflect'`` needs every pad smaller than the image side (one bounce, like ``F.pad``); OpenCV
    would keep bouncing -- the library call reports that case instead of guessing."""
    assert (shape is not None) ^ (padding is not None)
    if shape is not None:
        width = max(shape[1] - img.shape[1], 0)
        height = max(shape[0] - img.shape[0], 0)
        padding = (0, 0, width, height)
    if isinstance(pad_val, tuple):
        assert len(pad_val) == img.shape[-1]
    elif not isinstance(pad_val, numbers.Number):
        raise TypeError('pad_val must be a int or a tuple. '
                        f'But received {type(pad_val)}')
    if isinstance(padding, tuple) and len(padding) in [2, 4]:
        if len(padding) == 2:
            padding = (padding[0] - padding[0], padding[1], padding[0], padding[1])
    elif isinstance(padding, numbers.Number):
        padding = (padding, padding, padding, padding)
    else:
        raise ValueError('Padding must be a int or a 2, or 4 element tuple.'
                         f'But received {padding}')
    assert padding_mode in ['constant', 'edge', 'reflect', 'symmetric']
    mode = {'constant': 0, 'edge': 2, 'reflect': 1, 'symmetric': 4}[padding_mode]
    padding = tuple(int(v) for v in padding)
    top, bottom = padding[1] - padding[1] % 2, padding[3] + padding[1] % 2
    left, right = padding[0], padding[2]
    from .HexFrames import _Pad2dFn
    arr = np.asarray(img)
    chw = arr[None] if arr.ndim == 2 else np.transpose(arr, (2, 0, 1))
    bands = chw.shape[0]
    # border value as cv2.copyMakeBorder reads it (:722-730): a tuple is one value per band; a scalar is Scalar(v) =
    # (v, 0, 0, 0), so only band 0 of a multi-band image receives it -- and more than 4 bands need v == 0
    if isinstance(pad_val, tuple):
        vals = [float(v) for v in pad_val]
    else:
        if bands > 4 and float(pad_val) != 0.0:
            raise Exception("cv2.copyMakeBorder: a scalar border value must be 0 for images with more than 4 channels")
        vals = [float(pad_val)] + [0.0] * (bands - 1)
    if np.issubdtype(arr.dtype, np.integer):               # saturate_cast: round half to even, clamp to the type
        info = np.iinfo(arr.dtype)
        vals = [float(min(max(np.rint(v), info.min), info.max)) for v in vals]
    cast = None
    if chw.dtype not in (np.uint8, np.float32, np.float64):    # other integer types travel as float64 (exact up to 2^53)
        cast, chw = chw.dtype, chw.astype(np.float64)
    x = torch.from_numpy(np.ascontiguousarray(chw)).cuda()
    if mode != 0 or len(set(vals)) == 1:
        y = _Pad2dFn.apply(x, left, right, top, bottom, mode, vals[0])
    else:
        y = torch.cat([_Pad2dFn.apply(x[k:k + 1], left, right, top, bottom, mode, vals[k]) for k in range(bands)], 0)
    y = y.cpu().numpy()
    if cast is not None:
        y = y.astype(cast)
    # OpenCV drops a single-band axis: (H, W, 1) comes back as (H', W')
    return y[0] if bands == 1 else np.ascontiguousarray(np.transpose(y, (1, 2, 0)))


def hex_impad_to_multiple(img: np.ndarray, divisor: int, pad_val: Union[float, List] = 0) -> np.ndarray:
    """geometry_np.py:734-749."""
    pad_h = int(np.ceil(img.shape[0] / divisor)) * divisor
    pad_w = int(np.ceil(img.shape[1] / divisor)) * divisor
    return heximpad(img, shape=(pad_h, pad_w), pad_val=pad_val)
