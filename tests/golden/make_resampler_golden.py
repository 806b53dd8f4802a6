"""Golden vectors of the three learned lattice resamplers the reference retired into
``HyGrid/codes in old versions.txt``:
``Hex_to_Square_Conv2d_by_Double_Stride`` (lines 1-66), ``Square_to_Hex_Conv2d_by_Double_Stride`` (421-493) and
``Hex_to_Square_original_resolution`` (587-636), with the unfold helpers they call (637-739).

    python tests/golden/make_resampler_golden.py      # needs /root/reference (build container only)

The text file is not importable; the class bodies are exec'd UNMODIFIED in a namespace holding the names they expect from
the reference's own ``HexFrames`` module and from torch.  Output: tests/golden/resampler_golden.npz (inputs, perturbed
kernels, outputs, and the gradients autograd gives through the reference's own forward)."""
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, "/root/reference")
import HyGrid.HexFrames as hf   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resampler_golden.npz")
# (class, channels, even_odd_offset, factor, padding, padding_mode, H, W)
CASES = [("h2s", 3, 0, 2, 0, "constant", 8, 9), ("h2s", 2, 1, 2, 1, "constant", 7, 10), ("h2s", 4, 0, 4, 0, "constant", 12, 12),
         ("h2s", 2, 1, 4, 2, "constant", 13, 17), ("h2s", 1, 0, 6, 1, "constant", 19, 23), ("h2s", 3, 1, 2, 2, "reflect", 9, 8),
         ("s2h", 3, 0, 2, 0, "constant", 8, 9), ("s2h", 2, 0, 2, 1, "constant", 7, 10), ("s2h", 4, 0, 2, 2, "constant", 13, 17),
         ("s2h", 1, 0, 2, 0, "constant", 12, 12), ("s2h", 2, 0, 2, 1, "replicate", 10, 11),
         ("h2so", 3, 0, 2, 0, "constant", 8, 9), ("h2so", 2, 1, 2, 1, "constant", 7, 10), ("h2so", 4, 0, 2, 2, "constant", 12, 12),
         ("h2so", 1, 1, 2, 0, "constant", 3, 4), ("h2so", 2, 0, 2, 1, "reflect", 9, 9)]


def load_classes():
    from torch import Tensor
    from torch.nn import init
    src = open("/root/reference/HyGrid/codes in old versions.txt").read()
    ns = dict(nn=nn, torch=torch, np=np, F=F, math=math, init=init, Tensor=Tensor, pad=hf.pad,
              heximage_to_type1=hf.heximage_to_type1, type1_to_heximage=hf.type1_to_heximage)
    body = src[src.index("class Hex_to_Square_Conv2d_by_Double_Stride"):src.index("class HexPixelShuffle")]
    body += "\n" + src[src.index("class Square_to_Hex_Conv2d_by_Double_Stride"):src.index("class Quadtree_HexPooling")]
    body += "\n" + src[src.index("class Hex_to_Square_original_resolution"):]
    exec(compile(body, "codes in old versions.txt", "exec"), ns)
    return {"h2s": ns["Hex_to_Square_Conv2d_by_Double_Stride"], "s2h": ns["Square_to_Hex_Conv2d_by_Double_Stride"],
            "h2so": ns["Hex_to_Square_original_resolution"]}


def build(cls, kind, C, off, f, pad, mode):
    if kind == "h2s":
        return cls(C, off, f, padding=pad, padding_mode=mode)
    if kind == "s2h":
        return cls(C, f, padding=pad, padding_mode=mode)
    return cls(C, off, padding=pad, padding_mode=mode, trainable=True)


def main():
    classes = load_classes()
    torch.manual_seed(20261018)
    out = {"count": np.array(len(CASES))}
    for n, (kind, C, off, f, pad, mode, H, W) in enumerate(CASES):
        m = build(classes[kind], kind, C, off, f, pad, mode)
        out[f"{n}_init"] = m.kernel.detach().numpy().copy()           # the layer's own initial weights
        with torch.no_grad():
            m.kernel.add_(0.1 * torch.randn_like(m.kernel))
        x = torch.randn(2, C, H, W, requires_grad=True)
        y = m(x * 1.0)                                                # (* 1.0: h2so writes into its padded input)
        gy = torch.randn_like(y)
        (y * gy).sum().backward()
        out[f"{n}_cfg"] = np.array([{"h2s": 0, "s2h": 1, "h2so": 2}[kind], C, off, f, pad, {"constant": 0, "reflect": 1, "replicate": 2}[mode]])
        out[f"{n}_x"], out[f"{n}_kernel"], out[f"{n}_y"] = x.detach().numpy(), m.kernel.detach().numpy(), y.detach().numpy()
        out[f"{n}_gy"], out[f"{n}_dx"], out[f"{n}_dk"] = gy.numpy(), x.grad.numpy(), m.kernel.grad.numpy()
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
