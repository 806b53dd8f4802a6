#!/usr/bin/env python
"""Stress check (part of the test infrastructure, run by tools/gpu_suite.sh stress): tcgen05 wgrad vs the torch-CPU oracle for a list of shapes; prints where the error sits."""
import os, sys, itertools
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # tests/stress/ -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200"))
from HyGrid import HexFrames as hf
from oracle import hexframes_oracle as HO

def run(N, Cin, Cout, H, W, pad, off, pv=0.0):
    torch.manual_seed(3)
    xq = torch.randn(N, Cin, H, W).bfloat16().float()
    wq = (torch.randn(Cout, Cin, 1, 7) * 0.1).bfloat16().float()
    b = torch.randn(Cout)
    xr, wr = xq.clone().requires_grad_(), wq.clone().requires_grad_()
    ref = HO.hexconv2d(xr, wr, b, off, 2, 1, pad, 1, 1, padding_value=pv)
    gyq = torch.randn_like(ref).bfloat16().float()
    (ref * gyq).sum().backward()
    xg, wg, bg = xq.cuda().requires_grad_(), wq.cuda().requires_grad_(), b.cuda().requires_grad_()
    y = hf.hexconv2d(xg, wg, bg, off, 2, 1, pad, 1, 1, padding_value=pv, algo=2)
    (y * gyq.cuda()).sum().backward()
    d = (wg.grad.cpu() - wr.grad).abs()[:, :, 0, :]
    sc = float(wr.grad.abs().max())
    bad = d > 1e-3 * sc
    msg = f"N={N} Cin={Cin} Cout={Cout} H={H} W={W} pad={pad} off={off} pv={pv}: max err {float(d.max())/sc:.2e}"
    if bad.any():
        co = bad.any(2).any(1).nonzero().flatten().tolist()
        ci = bad.any(2).any(0).nonzero().flatten().tolist()
        k = bad.any(1).any(0).nonzero().flatten().tolist()
        msg += f"  BAD co={co[:8]}..({len(co)}) ci={ci[:8]}..({len(ci)}) k={k}"
    print(msg, flush=True)

if __name__ == "__main__":
    for cfg in [(1, 48, 16, 65, 64, 1, 1), (1, 48, 16, 65, 128, 1, 1), (1, 48, 16, 65, 64, 1, 0), (1, 48, 16, 32, 64, 1, 1),
                (1, 48, 16, 3, 64, 1, 1), (1, 48, 64, 65, 64, 1, 1), (1, 64, 16, 65, 64, 1, 1), (1, 32, 16, 65, 64, 1, 1),
                (1, 16, 16, 65, 64, 1, 1), (1, 48, 48, 65, 64, 1, 1), (1, 48, 32, 65, 64, 1, 1), (2, 48, 16, 65, 64, 1, 1),
                (1, 64, 64, 65, 64, 1, 1), (1, 64, 64, 64, 256, 1, 0)]:
        run(*cfg)
        run(*cfg, pv=0.25)
