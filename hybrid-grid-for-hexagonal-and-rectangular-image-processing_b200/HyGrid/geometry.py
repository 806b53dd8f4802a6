"""``HyGrid.geometry`` on B200: the names of the reference's legacy numba-CUDA module
(/root/reference/HyGrid/geometry.py), served by the same sm_100a kernels as ``geometry_np`` / ``geometry_torch``.

The reference module holds the repository's only hand-written GPU kernel (``resample_on_hexagonal_grids``,
geometry.py:8-152: one 32x32-thread launch per channel, fresh float64 host<->device copies of the coordinates and
the band every time).  Its arithmetic is that of the numpy twin -- float64 ``np.linspace`` coordinates, the same
affine-cell / triangle / area-weight rule -- up to one association: ``j_ = 0.5*i_ + (y_ + (w-.5)/2)`` (:28) against the
numpy twin's ``0.5*i_ + y_ + (w-.5)/2``, an ulp apart on some coordinates.  ``'linear'`` results here therefore agree with
it to 1e-12 of the value range (tests/golden/numba_twin_golden.npz, generated from the reference under
``NUMBA_ENABLE_CUDASIM=1``: every linspace case is bit-identical, one warp sample differs by 1.4e-14); on a real GPU
NVVM contracts the kernel's multiplies and adds into FMAs, so the reference is not bit-reproducible against itself.

What differs from ``geometry_np`` and is kept: results are always float64 (geometry.py:237,418); ``hex_to_square_resample``
returns ``(C, h1, w1)`` without ``.squeeze()`` (:435) while the warp squeezes (:262); the warp builds its output lattice
from ``np.mgrid[int(h_inf):h_sup+1, int(h_inf):h_sup+0.5]`` -- start truncated to an integer and the ROW extent reused
for the columns (:221).

Deviations: ``'nearest'`` resolves an exact three-way distance tie ``d1 == d3 < d2`` to the first candidate like
``geometry_torch`` does (the numba kernel's cascaded ``if``s pick the last one, :131-136; with NVVM's FMA contraction
the tie itself is not reproducible across GPUs); ``'bilinear'`` (a two-tap formula the reference itself marks as
unsupported, :52-87) raises NotImplementedError; ``hexresize`` works (the reference raises NameError on four undefined
names, :437-522) and follows ``geometry_np.hexresize``; the scipy ``griddata`` CPU variant is not provided.
"""
from __future__ import annotations

import numpy as np

from . import _native as nv
from . import functional as Fn
from ._hostapi import resample_host, to_device
from .geometry_np import _hex_method

__all__ = ["image_geometric_transformation_gpu", "image_geometric_transformation_cpu", "image_geometric_transformation",
           "hex_to_square_resample", "hexresize"]


def image_geometric_transformation_gpu(image: np.array, H: np.array = np.eye(3), interpolation='nearest', offset=0) -> np.array:
    """hex -> hex affine warp on the legacy lattice (geometry.py:156-262); float64 coordinates and result."""
    _hex_method(interpolation)
    out = Fn.hex_warp(to_device(image), H, interpolation, out_dtype=None, twin="numba")
    return out.cpu().numpy().astype(np.float64, copy=False).squeeze()


def image_geometric_transformation_cpu(image: np.array, H: np.array = np.eye(3), interpolation='nearest', offset=0) -> np.array:
    """geometry.py:264-345 is a scipy ``griddata`` CPU path; this build has no CPU path."""
    raise NotImplementedError("HyGrid on B200 has no CPU path")


def image_geometric_transformation(img: np.array, H: np.array = np.eye(3), interpolation='nearest', offset=0,
                                   device='cuda0') -> np.array:
    """Dispatcher (geometry.py:347-351): ``device`` in {'cuda0', 'cpu'}."""
    if device == 'cuda0':
        return image_geometric_transformation_gpu(img, H, interpolation, offset)
    if device == 'cpu':
        return image_geometric_transformation_cpu(img, H, interpolation, offset)


def hex_to_square_resample(hex_image, square_size=None, interpolation='nearest', offset=0):
    """geometry.py:354-435: float64 ``(C, h1, w1)``, not squeezed."""
    method = _hex_method(interpolation)
    out = resample_host("hex2rect", hex_image, square_size, method, np.float64, nv.MATH_EXACT, "np")
    return out.astype(np.float64, copy=False)


def hexresize(image, dsize, interpolation='linear'):
    """geometry.py:437-522 (unrunnable in the reference): the ``geometry_np.hexresize`` lattice, float64, not squeezed."""
    method = _hex_method(interpolation)
    out = resample_host("hexresize", image, dsize, method, np.float64, nv.MATH_EXACT, "np")
    return out.astype(np.float64, copy=False)
