#!/usr/bin/env python
"""Device-resident timings of every memory-bound kernel of the path at BASELINE sizes (configs 2 and 4):
achieved algorithmic GB/s (SURVEY.md 8d byte counts) against the measured HBM copy bandwidth.
CUDA events on the launching stream, 3 warm-ups, every tensor far larger than the 126 MB L2.

    python tools/bench_path.py [--reps 10] [--small]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200"))
from HyGrid import functional as Fn  # noqa: E402
from HyGrid import HexFrames as hf  # noqa: E402


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--only", default="", help="substring filter on the kernel names (profiling runs)")
    a = ap.parse_args()
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    dev = "cuda"
    rows = []

    def rec(name, ms, nbytes, pix):
        if callable(ms):
            if a.only and a.only not in name:
                return
            ms = ms()
        r = {"kernel": name, "ms": round(ms, 4), "alg_GB": round(nbytes / 1e9, 3), "GBps": round(nbytes / ms / 1e6, 1),
             "hbm_frac": round(nbytes / ms / 1e6 / hbm, 3), "hex_Mpix_s": round(pix / ms / 1e3, 1)}
        rows.append(r)
        print(json.dumps(r), flush=True)

    N4 = 8 if a.small else 64
    x4 = torch.rand(N4, 3, 2160, 3840, device=dev)
    planes, px = N4 * 3, 2160 * 3840
    y4 = torch.empty_like(x4)
    rec("c4 rect->hex bilinear fast f32", lambda: timeit(lambda: Fn.rect_to_hex(x4, None, "bilinear", out_dtype=torch.float32, math="fast", out=y4), a.reps), 8 * planes * px, N4 * px)
    for math in ("fast", "exact"):
        rec(f"c4 hex->rect linear {math} f32", lambda: timeit(lambda: Fn.hex_to_rect(x4, None, "linear", out_dtype=torch.float32, math=math, twin="np", out=y4), a.reps), 8 * planes * px, N4 * px)
    rec("c4 hex->rect nearest f32", lambda: timeit(lambda: Fn.hex_to_rect(x4, None, "nearest", twin="torch", out=y4), a.reps), 8 * planes * px, N4 * px)
    rec("c4 hexresize linear fast 2160x3840->1080x1920", lambda: timeit(lambda: Fn.hex_resize(x4, (1080, 1920), "linear", out_dtype=torch.float32, math="fast"), a.reps), 5 * planes * px, N4 * px)
    rec("c4 hex->type1 f32", lambda: timeit(lambda: Fn.hex_to_type1(x4, 0), a.reps), 4 * planes * (px + 2160 * (2 * 3840 + 1)), N4 * px)
    # pooling pyramid (config 4): average 2x2, five levels
    pool = hf.HexPool2d("average", 2, 2)
    cur = x4
    lvl_in = []
    for lvl in range(5):
        lvl_in.append(cur)
        cur = pool(cur)
    for lvl, t in enumerate(lvl_in[:3]):
        hn, wn = (t.shape[-2] - 2) // 2 + 1, (t.shape[-1] - 1) // 2
        rec(f"c4 pool avg 2x2 level {lvl} {t.shape[-2]}x{t.shape[-1]}", lambda: timeit(lambda: pool(t), a.reps),
            4 * planes * (t.shape[-2] * t.shape[-1] + hn * wn), N4 * t.shape[-2] * t.shape[-1])

    def pyramid():
        c = x4
        for _ in range(5):
            c = pool(c)
        return c
    rec("c4 pool pyramid (5 levels)", lambda: timeit(pyramid, a.reps), 4 * planes * px * (1 + 0.25) * 1.34, N4 * px)
    mp = hf.HexPool2d("max", 2, 2)
    xg = x4.requires_grad_()
    yp = mp(xg)
    gy = torch.rand_like(yp)
    rec("c4 pool max fwd(+aux)", lambda: timeit(lambda: mp(xg), a.reps), 4 * planes * (px + yp.shape[-2] * yp.shape[-1]), N4 * px)
    rec("c4 pool max bwd", lambda: timeit(lambda: torch.autograd.grad(mp(xg), xg, gy), max(2, a.reps // 2)) - timeit(lambda: mp(xg), a.reps), 4 * planes * (px + yp.shape[-2] * yp.shape[-1]), N4 * px)
    del x4, y4, lvl_in, cur, xg, yp, gy
    torch.cuda.empty_cache()

    N2 = 32 if a.small else 256
    x2 = torch.rand(N2, 3, 1024, 1024, device=dev) * 255
    y2 = torch.empty_like(x2)
    p2, q2 = N2 * 3, 1024 * 1024
    rec("c2 rect->hex bilinear fast", lambda: timeit(lambda: Fn.rect_to_hex(x2, None, "bilinear", out_dtype=torch.float32, math="fast", out=y2), a.reps), 8 * p2 * q2, N2 * q2)
    rec("c2 rect->hex nearest f32", lambda: timeit(lambda: Fn.rect_to_hex(x2, None, "nearest", out=y2), a.reps), 8 * p2 * q2, N2 * q2)
    u8 = (x2[: N2 // 2]).to(torch.uint8)
    rec("c1-style rect->hex nearest u8 -> half res", lambda: timeit(lambda: Fn.rect_to_hex(u8, (512, 512), "nearest"), a.reps), (N2 // 2) * 3 * (q2 // 4 * 2), (N2 // 2) * q2 // 4)
    rec("c2 hex->rect linear fast", lambda: timeit(lambda: Fn.hex_to_rect(x2, None, "linear", out_dtype=torch.float32, math="fast", twin="np", out=y2), a.reps), 8 * p2 * q2, N2 * q2)
    del x2, y2
    torch.cuda.empty_cache()

    # R3: hex -> hex affine warp with the inverse map evaluated in-kernel (2 degree rotation + shift), float32 coordinates
    import numpy as np
    Nw = 4 if a.small else 32
    xw = torch.rand(Nw, 3, 2160, 3840, device=dev)
    th = np.deg2rad(2.0)
    Hm = np.array([[np.cos(th), -np.sin(th), 3.0], [np.sin(th), np.cos(th), -5.0], [0, 0, 1.0]])
    for interp in ("linear", "nearest"):
        yw = Fn.hex_warp_affine(xw, Hm, interp)
        rec(f"r3 hex warp affine {interp} 2160x3840", lambda: timeit(lambda: Fn.hex_warp_affine(xw, Hm, interp), a.reps),
            4 * (xw.numel() + yw.numel()), Nw * yw.shape[-2] * yw.shape[-1])
    del xw, yw
    torch.cuda.empty_cache()

    # table-driven gathers
    Ns = 4 if a.small else 32
    xs_ = torch.randn(Ns, 64 * 4, 256, 256, device=dev)                      # 64 output channels, upscale 2
    ps = hf.HexPixelShuffle(2)
    ys_ = ps(xs_)
    rec("pixel shuffle r=2 64ch 256x256 f32", lambda: timeit(lambda: ps(xs_), a.reps), 4 * (ys_.numel() * 2), Ns * ys_.shape[2] * ys_.shape[3])
    del xs_, ys_
    from HyGrid.HexPixelArt import hexagon_mosaic
    xm = torch.randint(0, 256, (Ns * 3, 1024, 1024), device=dev, dtype=torch.uint8)
    rec("hex mosaic 1024x1024 u8 -> 4096x4096", lambda: timeit(lambda: hexagon_mosaic(xm, (4096, 4096)), a.reps),
        xm.numel() + xm.shape[0] * 4096 * 4096, Ns * 1024 * 1024)
    print(json.dumps({"hbm_peak_gbs": hbm, "rows": rows}))


if __name__ == "__main__":
    main()
