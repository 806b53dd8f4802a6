#!/usr/bin/env python
"""Stress check (test infrastructure; run on a GPU box: python tests/stress/conv_narrow_fuzz.py [cases] [seed]): the tcgen05 conv
kernels on lattices narrower than a 128-pixel tile with MORE work items than CTAs -- compact ring slots, loader warp groups,
x / gy converter groups, ceil(W / 16) reduction steps, 128-channel passes, TMA boxes past the channel count (RGB) -- forward, data,
weight and bias gradient against the torch-CPU oracle on bf16-exact operands.  Prints one line per case; BAD marks a failure."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # tests/stress/ -> repo root
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200"))
from HyGrid import HexFrames as hf  # noqa: E402
from oracle import hexframes_oracle as HO  # noqa: E402


def run(N, Cin, Cout, H, W, pad, off, pv):
    torch.manual_seed(Cin * 131 + Cout * 17 + H * 3 + W)
    xq = torch.randn(N, Cin, H, W).bfloat16().float()
    wq = (torch.randn(Cout, Cin, 1, 7) * 0.1).bfloat16().float()
    b = torch.randn(Cout)
    xr, wr, br = xq.clone().requires_grad_(), wq.clone().requires_grad_(), b.clone().requires_grad_()
    try:
        ref = HO.hexconv2d(xr, wr, br, off, 2, 1, pad, 1, 1, padding_value=pv)
    except Exception:
        return None
    gyq = torch.randn_like(ref).bfloat16().float()
    (ref * gyq).sum().backward()
    xg, wg, bg = xq.cuda().requires_grad_(), wq.cuda().requires_grad_(), b.cuda().requires_grad_()
    y = hf.hexconv2d(xg, wg, bg, off, 2, 1, pad, 1, 1, padding_value=pv, algo=2)
    (y * gyq.cuda()).sum().backward()
    errs = {"y": float((y.detach().cpu() - ref.detach()).abs().max()) / float(ref.detach().abs().max()),
            "dx": float((xg.grad.cpu() - xr.grad).abs().max()) / float(xr.grad.abs().max()),
            "dw": float((wg.grad.cpu() - wr.grad).abs().max()) / float(wr.grad.abs().max()),
            "db": float((bg.grad.cpu() - br.grad).abs().max()) / max(float(br.grad.abs().max()), 1.0)}
    bad = errs["y"] > 1e-4 or errs["dx"] > 1e-4 or errs["dw"] > 1e-3 or errs["db"] > 1e-3
    print(f"N={N} Cin={Cin} Cout={Cout} H={H} W={W} pad={pad} off={off} pv={pv}: " + " ".join(f"{k} {v:.1e}" for k, v in errs.items())
          + ("  BAD" if bad else ""), flush=True)
    return bad


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
    nbad = 0
    for _ in range(n):
        cin = int(rng.choice([16, 32, 48, 64, 96, 128]))
        cout = int(rng.choice([16, 32, 64, 128]))
        W = int(rng.integers(8, 131))
        if rng.random() < 0.4:
            W = W // 4 * 4 + 4                     # 16-byte rows: the TMA variant
        cfg = (int(rng.integers(8, 49)), cin, cout, int(rng.integers(6, 41)), W, int(rng.integers(0, 3)), int(rng.integers(0, 2)),
               0.0 if rng.random() < 0.8 else 0.25)
        nbad += bool(run(*cfg))
    print(f"{nbad} BAD of {n}")
