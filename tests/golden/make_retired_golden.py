"""Golden vectors of ops the reference retired into ``HyGrid/codes in old versions.txt``.

    python tests/golden/make_retired_golden.py      # needs /root/reference (build container only)

The text file is not importable; the class bodies of ``HexPixelShuffle`` (lines 68-126) and ``HexConvTranspose2d``
(lines 129-274) are exec'd UNMODIFIED in a namespace holding the names they expect from the reference's own
``HexFrames`` module and from torch.  Output:
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


# (radius, stride, even_odd_offset, groups, Cin, Cout, H, W, bias)
CT_CASES = [(2, 1, 0, 1, 4, 6, 6, 7, True), (2, 2, 0, 1, 4, 6, 6, 7, True), (2, 2, 1, 2, 4, 6, 5, 5, False),
            (3, 2, 0, 1, 2, 3, 8, 6, True), (2, 3, 1, 1, 3, 5, 6, 7, True), (3, 3, 0, 1, 2, 2, 4, 9, False),
            (2, 2, 1, 1, 16, 16, 9, 12, True), (3, 1, 1, 2, 4, 4, 5, 5, True)]


def load_class():
    from torch import Tensor
    from torch.nn import init
    src = open("/root/reference/HyGrid/codes in old versions.txt").read()
    ns = dict(nn=nn, torch=torch, np=np, F=F, math=math, init=init, Tensor=Tensor, pad=hf.pad,
              heximage_to_type1=hf.heximage_to_type1, type1_to_heximage=hf.type1_to_heximage)
    body = src[src.index("class HexPixelShuffle"):src.index("class im2col_HexConv2d")]
    exec(compile(body, "codes in old versions.txt[68:274]", "exec"), ns)
    return ns["HexPixelShuffle"], ns["HexConvTranspose2d"]


def main():
    PS, CT = load_class()
    rng = np.random.default_rng(20260319)
    out = {"count": np.array(len(CASES))}
    for n, (r, cout, H, W) in enumerate(CASES):
        x = rng.integers(-99, 100, (2, cout * r * r, H, W)).astype(np.float32)
        y = PS(r)(torch.from_numpy(x))
        out[f"ps_{n}_r"] = np.array(r)
        out[f"ps_{n}_in"] = x
        out[f"ps_{n}_out"] = y.numpy()
    torch.manual_seed(20260319)
    out["ct_count"] = np.array(len(CT_CASES))
    for n, (r, st, eo, g, cin, cout, H, W, bias) in enumerate(CT_CASES):
        m = CT(cin, cout, eo, r, stride=st, groups=g, bias=bias)
        x = torch.randn(2, cin, H, W)
        with torch.no_grad():
            y = m(x)
        out[f"ct_{n}_cfg"] = np.array([r, st, eo, g, cin, cout, int(bias)])
        out[f"ct_{n}_in"] = x.numpy()
        out[f"ct_{n}_kernel"] = m.kernel.detach().numpy()
        out[f"ct_{n}_bias"] = m.bias.detach().numpy() if bias else np.zeros(0, np.float32)
        out[f"ct_{n}_out"] = y.numpy()
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
