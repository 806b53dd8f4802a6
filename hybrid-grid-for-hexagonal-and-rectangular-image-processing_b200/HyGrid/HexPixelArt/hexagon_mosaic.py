"""Rect preview of a hex image: the fragment shader of the reference's viewer as a table-driven gather.

Reference: HexPixelArt/hexagon_mosaic_shader.py:25-81 (fragment shader), texture.py:31-40 (the texture is padded with
zero rows / columns to multiples of 4; ``size`` is the PADDED size, :61), HexImage.py:230-235 + shader :19-20 (a
full-window quad whose texture coordinate runs left -> right and TOP -> bottom).

Per output pixel the shader (float32) scales ``uv`` to lattice units (``x = u*(size.x+0.5)``, ``y = v*(size.y+1)``),
finds the half-cell box ``(wx, wy) = (int(x/TB), int(y/TR))`` (``TR = 2^-hierarchy``, ``TB = TR/2``), takes the two
box corners that are cell centres for that box parity (:51-58), keeps the nearer one (:60-69, ties -> the second) and
turns it into texel ``row = vy - 1``, ``col = (vx - 1 - (vy+1+even_odd_offset)%2) / 2`` (integer division, :71-75).
The sample point is a texel centre, so ``GL_LINEAR`` returns that texel itself; outside the texture
``GL_CLAMP_TO_BORDER`` gives black.  The rule depends on shapes only -> one int64 table per
(H, W, out size, offset, hierarchy), built on the host in float32 like the shader, then ``hg_plane_gather``.

Not reproduced: mip-mapped minification (texture.py:50: the hardware picks a mip level from screen-space derivatives of
a piecewise-constant coordinate, i.e. blurs pixels on cell borders when the window is smaller than the texture) and the
1-band -> RGB replication of texture.py:27-28 (bands stay as they are).  **Parity unpinned**: GLSL cannot run in the
build container (no OpenGL); ``oracle/hexmosaic_oracle.py`` restates the shader statement by statement.
"""
from __future__ import annotations

import functools

import numpy as np
import torch

from .. import _native as nv

_F = np.float32


@functools.lru_cache(maxsize=32)
def mosaic_table(H: int, W: int, out_h: int, out_w: int, even_odd_offset: int = 0, hierarchy: int = 0):
    """int64 ``(out_h, out_w)`` array of source offsets ``row*W + col`` into an ``H x W`` hex image, -1 = border."""
    th, tw = -(-H // 4) * 4, -(-W // 4) * 4                      # texture.py:31-40
    eo = int(even_odd_offset)
    TR = _F(2.0 ** (-int(hierarchy)))
    TB = _F(0.5) * TR
    x = ((np.arange(out_w, dtype=_F) + _F(0.5)) / _F(out_w)) * (_F(tw) + _F(0.5))
    y = ((np.arange(out_h, dtype=_F) + _F(0.5)) / _F(out_h)) * (_F(th) + _F(1))
    wx = (x / TB).astype(np.int32)[None, :]                      # int(): truncation
    wy = (y / TR).astype(np.int32)[:, None]
    x, y = x[None, :], y[:, None]
    same = ((wx + eo) & 1) == (wy & 1)
    v1x, v2x = TB * wx.astype(_F), TB * (wx.astype(_F) + _F(1))
    v1y = np.where(same, TR * wy.astype(_F), TR * (wy.astype(_F) + _F(1)))
    v2y = np.where(same, TR * (wy.astype(_F) + _F(1)), TR * wy.astype(_F))
    s1 = (v1x - x) * (v1x - x) + (v1y - y) * (v1y - y)
    s2 = (v2x - x) * (v2x - x) + (v2y - y) * (v2y - y)
    first = s1 < s2
    vx = (np.where(first, v1x, v2x) / _F(0.5)).astype(np.int32)
    vy = np.where(first, v1y, v2y).astype(np.int32)
    num = vx - 1 - (vy + 1 + eo) % 2
    col = np.where(num >= 0, num // 2, -((-num) // 2))            # GLSL / C integer division truncates toward zero
    row = vy - 1
    inside = (row >= 0) & (row < H) & (col >= 0) & (col < W)     # the zero padding up to (th, tw) is black as well
    return np.where(inside, row.astype(np.int64) * W + col, -1)


@functools.lru_cache(maxsize=32)
def _device_table(H, W, out_h, out_w, eo, hierarchy, device):
    tab = mosaic_table(H, W, out_h, out_w, eo, hierarchy)
    return torch.from_numpy(np.ascontiguousarray(tab.reshape(-1))).to(torch.device(device))


def hexagon_mosaic(hex_image, out_size=None, even_odd_offset: int = 0, hierarchy: int = 0):
    """Hex-mosaic raster of ``hex_image`` (``(..., H, W)``; numpy array or CUDA tensor; uint8 / float32 / float64 /
    bfloat16): every output pixel shows the hex cell it falls in, black outside the lattice.  ``out_size`` defaults to
    4 pixels per texel of the padded texture.  ``hierarchy`` is the viewer's mosaic level (cell size 2^-hierarchy).
    numpy in -> numpy out (same dtype); tensor in -> tensor out on the same device."""
    is_np = not isinstance(hex_image, torch.Tensor)
    if is_np:
        arr = np.asarray(hex_image)
        if arr.ndim < 2:
            raise Exception(f"dim of image should be 2 or more, but got dim = {arr.ndim} instead")
        cast = None
        if arr.dtype not in (np.uint8, np.float32, np.float64):
            cast, arr = arr.dtype, arr.astype(np.float64)
        if not torch.cuda.is_available():
            raise nv.HyGridNativeError("no CUDA device: HyGrid on B200 has no CPU path")
        x = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    else:
        x = nv.require_cuda(hex_image, "hex_image")
        if x.dim() < 2:
            raise Exception(f"dim of image should be 2 or more, but got dim = {x.dim()} instead")
        if x.dtype not in (torch.uint8, torch.float32, torch.float64, torch.bfloat16):
            raise TypeError(f"unsupported tensor dtype {x.dtype}")
        x = x.contiguous()
    H, W = int(x.shape[-2]), int(x.shape[-1])
    if out_size is None:
        out_size = (4 * (-(-H // 4) * 4), 4 * (-(-W // 4) * 4))
    out_h, out_w = int(out_size[0]), int(out_size[1])
    if out_h <= 0 or out_w <= 0 or H <= 0 or W <= 0:
        raise ValueError(f"empty image or raster: image {H}x{W}, raster {out_h}x{out_w}")
    planes = x.numel() // (H * W)
    table = _device_table(H, W, out_h, out_w, int(even_odd_offset), int(hierarchy), str(x.device))
    y = torch.empty(x.shape[:-2] + (out_h, out_w), dtype=x.dtype, device=x.device)
    dt = nv.hg_dtype(x.dtype)
    nv.call("hg_plane_gather", nv.ptr(x), nv.ptr(y), nv.ptr(table), 1, planes, out_h * out_w, planes * H * W, H * W,
            dt, dt, nv.stream_ptr(x.device))
    if not is_np:
        return y
    res = y.cpu().numpy()
    return res.astype(cast) if cast is not None else res
