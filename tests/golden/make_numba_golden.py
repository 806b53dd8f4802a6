"""Golden vectors of the reference's legacy numba twin (HyGrid/geometry.py), run UNMODIFIED in this container.

    python tests/golden/make_numba_golden.py        # needs /root/reference (build container only)

There is no GPU here, so numba's own simulator executes the ``@cuda.jit`` kernel (``NUMBA_ENABLE_CUDASIM=1``: plain
IEEE double arithmetic, no FMA contraction).  ``torch.cuda.empty_cache`` -- called once per channel by the reference
(geometry.py:241,422) -- is a no-op without a device and is stubbed as one.  Cases are small because the simulator
runs every CUDA thread as a Python thread.  Output: tests/golden/numba_twin_golden.npz.
"""
import os
import sys

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
sys.path.insert(0, "/root/reference")

import numpy as np
import torch

torch.cuda.empty_cache = lambda: None
from HyGrid import geometry as G   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "numba_twin_golden.npz")

# (name, C, h, w, h1, w1): 'nearest' cases avoid an exact d1 == d3 tie (see HyGrid/geometry.py docstring)
RESAMPLE = [("up2", 2, 12, 10, 24, 20), ("odd", 1, 9, 7, 13, 17), ("same", 3, 8, 8, 8, 8), ("down", 2, 14, 18, 9, 11),
            ("band2d", 0, 10, 12, 15, 9)]
t = np.deg2rad(20.0)
WARP = [("identity", 2, 9, 8, np.eye(3)),
        ("scale", 1, 8, 10, np.array([[1.5, 0, 0], [0, 0.75, 0], [0, 0, 1.0]])),
        ("rotate", 2, 10, 9, np.array([[np.cos(t), -np.sin(t), 0.3], [np.sin(t), np.cos(t), -0.6], [0, 0, 1.0]]))]


def main():
    rng = np.random.default_rng(20260318)
    out = {}
    for name, c, h, w, h1, w1 in RESAMPLE:
        img = rng.random((c, h, w) if c else (h, w)) * 255
        out[f"resample/{name}/in"] = img
        out[f"resample/{name}/size"] = np.array([h1, w1])
        for interp in ("linear", "nearest"):
            out[f"resample/{name}/{interp}"] = G.hex_to_square_resample(img, (h1, w1), interp)
    for name, c, h, w, H in WARP:
        img = rng.random((c, h, w)) * 255
        out[f"warp/{name}/in"] = img
        out[f"warp/{name}/H"] = H
        for interp in ("linear", "nearest"):
            out[f"warp/{name}/{interp}"] = G.image_geometric_transformation_gpu(img, H, interp)
    np.savez_compressed(OUT, **out)
    print(OUT, len(out), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
