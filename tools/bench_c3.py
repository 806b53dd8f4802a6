#!/usr/bin/env python
"""BASELINE config 3: HexModules hex-conv stack -- 4 x HexConv2d(64, 64, radius 2, padding 1) on a 256 x 256 hex
lattice, batch 128, autocast bfloat16, loss = y.float().sum(), forward + backward.  CUDA events, 3
warm-ups.  hex Mpix/s = N * H * W per fwd+bwd step; per layer-step = that x layers."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200"))
from HyGrid import HexFrames as hf  # noqa: E402
from HyGrid import _native as nv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[hf.HexConv2d(64, 64, 0, 2, stride=1, padding=1) for _ in range(a.layers)]).cuda()
    x = torch.randn(a.batch, 64, 256, 256, device="cuda")

    def step():
        for p in net.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = net(x)
        y.float().sum().backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    nv.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    pix = a.batch * 256 * 256
    flops = a.layers * 3 * 2 * 7 * 64 * 64 * pix          # fwd + dgrad + wgrad (the first layer's dgrad is not needed: minus one)
    flops -= 2 * 7 * 64 * 64 * pix
    print(json.dumps({"workload": f"C3: {a.layers} x HexConv2d(64,64,r=2) on 256x256, batch {a.batch}, autocast bf16, fwd+bwd",
                      "ms_per_step": round(ms, 3), "hex_mpix_per_s": round(pix / ms / 1e3, 1),
                      "hex_mpix_per_s_per_layer_step": round(pix * a.layers / ms / 1e3, 1), "tflops": round(flops / ms / 1e9, 1),
                      "library_launches_per_step": nv.launch_count() // a.reps}))


if __name__ == "__main__":
    main()
