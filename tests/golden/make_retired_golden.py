"""Golden vectors of ops the reference retired into ``HyGrid/codes in old versions.txt``.

    python tests/golden/make_retired_golden.py      # needs /root/reference (build container only)

The text file is not importable; the class body of ``HexPixelShuffle`` (lines 68-126) is exec'd UNMODIFIED in a
namespace holding the names it expects from the reference's own ``HexFrames`` module.  Output:
tests/golden/retired_golden.npz (inputs are seeded integers stored as float32, outputs the class's results).
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, "/root/reference")
import HyGrid.HexFrames as hf   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "retired_golden.npz")
CASES = [(2, 1, 4, 5), (2, 2, 5, 4), (2, 1, 1, 3), (3, 1, 4, 5), (3, 2, 3, 3), (3, 1, 7, 2), (4, 1, 6, 6), (4, 2, 3, 5),
         (5, 1, 4, 3), (2, 3, 8, 8)]     # (upscale_factor, out_channels, H, W)


def load_class():
    src = open("/root/reference/HyGrid/codes in old versions.txt").read()
    body = src[src.index("class HexPixelShuffle"):src.index("class HexConvTranspose2d")]
    ns = dict(nn=nn, torch=torch, np=np, F=F, math=math, pad=hf.pad, heximage_to_type1=hf.heximage_to_type1,
              type1_to_heximage=hf.type1_to_heximage)
    exec(compile(body, "codes in old versions.txt[68:126]", "exec"), ns)
    return ns["HexPixelShuffle"]


def main():
    PS = load_class()
    rng = np.random.default_rng(20260319)
    out = {"count": np.array(len(CASES))}
    for n, (r, cout, H, W) in enumerate(CASES):
        x = rng.integers(-99, 100, (2, cout * r * r, H, W)).astype(np.float32)
        y = PS(r)(torch.from_numpy(x))
        out[f"ps_{n}_r"] = np.array(r)
        out[f"ps_{n}_in"] = x
        out[f"ps_{n}_out"] = y.numpy()
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
