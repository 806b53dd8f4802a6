"""CPU oracle for the hex-mosaic preview: the viewer's GLSL fragment shader, one pixel at a time.

TEST INFRASTRUCTURE ONLY (see the header of ``hygrid_oracle.py`` for who may import this).

**Pin.**  The reference's implementation of this step is a GLSL fragment shader
(HyGrid/HexPixelArt/hexagon_mosaic_shader.py:25-81) that only runs inside an OpenGL context; the build container has
no OpenGL / GLFW.  The shader's own source text was therefore executed on the CPU: tests/golden/make_mosaic_golden.py cuts it
out of the reference file, tests/golden/glsl_mini.py translates it statement by statement (binary32 floats, truncating
integer division / conversion, implicit int -> float) and the texel coordinates it produces are committed as
tests/golden/mosaic_golden.npz; tests/test_zz_hex_mosaic.py checks this restatement against them fragment by fragment.
The GL pipeline around the shader (texture filtering, mip-mapping) stays unpinned.  What follows restates the shader statement
by statement in float32 (GLSL ``float``), with ``int()`` as truncation and ``/`` on ints as C division; texture
filtering is reduced to what it evaluates to at a texel centre (the texel itself; black outside,
``GL_CLAMP_TO_BORDER``, texture.py:47-48).  Mip-mapped minification (texture.py:50) is not modelled.

All ``file:line`` citations are into ``/root/reference/HyGrid/HexPixelArt/``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["fragment_cell", "hexagon_mosaic"]

f32 = np.float32


def _cdiv(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def fragment_cell(u, v, size_x, size_y, even_odd_offset=0, ratio=1.0):
    """hexagon_mosaic_shader.py:38-76 for one fragment: texture coordinate (u, v) -> texel (row, col)."""
    u, v, ratio = f32(u), f32(v), f32(ratio)
    sizex = f32(size_x) + f32(0.5)                    # :40
    sizey = f32(size_y) + f32(1)                      # :42
    TR = f32(1) * ratio                               # :43
    TB = f32(0.5) * ratio                             # :44
    x = u * sizex                                     # :46
    y = v * sizey                                     # :47
    wx = int(x / TB)                                  # :49
    wy = int(y / TR)                                  # :50
    if ((wx + even_odd_offset) & 1) == (wy & 1):      # :54-57
        v1 = (TB * f32(wx), TR * f32(wy))
        v2 = (TB * (f32(wx) + f32(1)), TR * (f32(wy) + f32(1)))
    else:                                             # :58-61
        v1 = (TB * f32(wx), TR * (f32(wy) + f32(1)))
        v2 = (TB * (f32(wx) + f32(1)), TR * f32(wy))
    s1 = (v1[0] - x) * (v1[0] - x) + (v1[1] - y) * (v1[1] - y)   # :63
    s2 = (v2[0] - x) * (v2[0] - x) + (v2[1] - y) * (v2[1] - y)   # :64
    vv = v1 if s1 < s2 else v2                        # :67-72
    vx = int(vv[0] / f32(0.5))                        # :74
    vy = int(vv[1] / f32(1))                          # :75
    sx_int = _cdiv(vx - 1 - (vy + 1 + even_odd_offset) % 2, 2)   # :77 (integer arithmetic up to the "+0.5")
    # sx = sx_int + 0.5, sy = vy - 0.5 (:77-78): the centre of texel (row vy-1, col sx_int)
    return vy - 1, sx_int


def hexagon_mosaic(hex_image, out_size, even_odd_offset=0, hierarchy=0):
    """(C, H, W) hex image -> (C, out_h, out_w) raster, same dtype.  The texture is the image zero-padded to multiples
    of 4 (texture.py:31-40) and ``size`` is that padded size (:61); fragment centres are at ((px+.5)/out_w,
    (py+.5)/out_h) with v running top -> bottom (shader :19-20 flips aTex.y, HexImage.py:230-235)."""
    img = np.asarray(hex_image)
    C, H, W = img.shape
    th, tw = (H + 3) // 4 * 4, (W + 3) // 4 * 4
    out_h, out_w = out_size
    out = np.zeros((C, out_h, out_w), img.dtype)
    ratio = 2.0 ** (-hierarchy)                        # texture.py:62
    for py in range(out_h):
        v = (f32(py) + f32(0.5)) / f32(out_h)
        for px in range(out_w):
            u = (f32(px) + f32(0.5)) / f32(out_w)
            row, col = fragment_cell(u, v, tw, th, even_odd_offset, ratio)
            if 0 <= row < H and 0 <= col < W:
                out[:, py, px] = img[:, row, col]
    return out
