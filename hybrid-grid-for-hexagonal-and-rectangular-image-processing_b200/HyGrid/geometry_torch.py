"""``HyGrid.geometry_torch`` on B200 (/root/reference/HyGrid/geometry_torch.py): numpy in, numpy out,
computed on the GPU.  The reference does the coordinate arithmetic on CPU tensors, ships eight index
tensors to the device, gathers there and ships flags back; here the only host work is the 1-D
``torch.linspace`` tables (float32 then widened, exactly geometry_torch.py:252-253 -- that float32
rounding is part of the reference's results) and everything else is one kernel launch.
"""
from __future__ import annotations

import numpy as np

from . import _native as nv
from . import functional as Fn
from ._hostapi import as_chw, resample_host, to_device
from .geometry_np import _hex_method

__all__ = ["image_geometric_transformation_gpu", "hex_to_square_resample", "image_geometric_transformation_cpu",
           "image_geometric_transformation"]


def image_geometric_transformation_gpu(image: np.array, H: np.array = np.eye(3), interpolation='nearest', offset=0) -> np.array:
    """hex -> hex affine warp, float32 inverse-mapped coordinates (geometry_torch.py:7-189)."""
    _hex_method(interpolation)
    x = to_device(image)
    out = Fn.hex_warp(x, H, interpolation, twin="torch")
    res = out.cpu().numpy()
    if interpolation == 'nearest':
        res = res.astype(as_chw(image).dtype, copy=False)
    return res.squeeze()


def hex_to_square_resample(hex_image, square_size=None, interpolation='nearest', offset=0):
    """geometry_torch.py:191-358."""
    method = _hex_method(interpolation)
    return resample_host("hex2rect", hex_image, square_size, method, np.float64, nv.MATH_EXACT, "torch").squeeze()


def image_geometric_transformation_cpu(img: np.array, H: np.array = np.eye(3), interpolation='nearest', offset=0) -> np.array:
    """The reference's scipy ``griddata`` path is broken (undefined name, geometry_torch.py:366) and is a CPU
    path; this build has none."""
    raise NotImplementedError("HyGrid on B200 has no CPU path (the reference's CPU variant raises NameError)")


def image_geometric_transformation(img: np.array, H: np.array = np.eye(3), interpolation='nearest', offset=0,
                                   device='cuda0') -> np.array:
    """Dispatcher (geometry_torch.py:442-446): ``device`` in {'cuda0', 'cpu'}."""
    if device == 'cuda0':
        return image_geometric_transformation_gpu(img, H, interpolation, offset)
    if device == 'cpu':
        return image_geometric_transformation_cpu(img, H, interpolation, offset)
