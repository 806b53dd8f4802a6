"""Fixtures for the hex-mosaic preview from the REFERENCE'S OWN fragment shader text.

    python tests/golden/make_mosaic_golden.py        # needs /root/reference (build container only)

The viewer's pixel -> hex-cell rule is a GLSL fragment shader (HyGrid/HexPixelArt/hexagon_mosaic_shader.py:26-82) and the
build container has no OpenGL.  The shader source is cut out of the reference file as text, translated statement by statement
by tests/golden/glsl_mini.py (GLSL value semantics: binary32 floats, truncating int division / conversion, implicit
int -> float) and executed for every fragment centre of a few rasters.  Stored: the texel coordinate (sx, sy) the shader hands
to ``texture2D``, per fragment.  Output: tests/golden/mosaic_golden.npz."""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from glsl_mini import compile_fragment_shader  # noqa: E402

OUT = os.path.join(HERE, "mosaic_golden.npz")
# (texture rows, texture cols [multiples of 4, texture.py:31-40], raster h, raster w, even_odd_offset, hierarchy)
CASES = [(8, 8, 24, 20, 0, 0), (8, 12, 31, 40, 1, 0), (4, 4, 16, 16, 0, 0), (12, 8, 50, 37, 1, 1), (8, 8, 64, 64, 0, 2), (16, 20, 45, 70, 1, 0)]


def main():
    text = open("/root/reference/HyGrid/HexPixelArt/hexagon_mosaic_shader.py", encoding="utf-8").read()
    fs = re.search(r'fs\s*=\s*"""(.*?)"""', text, flags=re.S).group(1)
    run, code = compile_fragment_shader(fs)
    out = {"count": np.array(len(CASES)), "translated_source": np.array(code)}
    for n, (th, tw, oh, ow, eoo, hier) in enumerate(CASES):
        sx = np.zeros((oh, ow), np.float32)
        sy = np.zeros((oh, ow), np.float32)
        ratio = 2.0 ** (-hier)                                   # texture.py:62
        for py in range(oh):
            v = (np.float32(py) + np.float32(0.5)) / np.float32(oh)      # fragment centre; aTex.y is flipped by the vertex shader (:19-20)
            for px in range(ow):
                u = (np.float32(px) + np.float32(0.5)) / np.float32(ow)
                a, b = run(u, v, tw, th, ratio, eoo)             # size = (texture width, texture height), texture.py:61
                sx[py, px], sy[py, px] = a.v, b.v
        out[f"{n}_cfg"] = np.array([th, tw, oh, ow, eoo, hier])
        out[f"{n}_sx"], out[f"{n}_sy"] = sx, sy
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
