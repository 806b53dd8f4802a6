#!/usr/bin/env python
"""A/B sweeps of the resampling kernels' run-time switches at BASELINE sizes, in ONE process (the switches are read on
every call): which variant ships as the default is decided from this file's output, committed under profiles/.

    python tools/sweep_kernels.py [--what r2h,h2r] [--reps 10] > gpurun_out/<tag>/sweep_kernels.jsonl

rect->hex bilinear (C2 256x3x1024^2, C4 64x3x2160x3840): the row-streaming kernel (HG_R2H_STREAM_ROWS x HG_R2H_STREAM_PF)
against the TMA warp-specialised kernel (HG_R2H_STREAM=0), float32 fast / exact, and exact with a float64 result.
hex->rect linear: plane-group sharing HG_HEXSRC_SHARE x warps per CTA HG_HEXSRC_WARPS, fast and exact.
CUDA events on the launching stream, 3 warm-ups, every tensor far larger than the 126 MB L2."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200"))
from HyGrid import _native as nv  # noqa: E402
from HyGrid import functional as Fn  # noqa: E402


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


class Env:
    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items() if v is not None}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="r2h,h2r")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0

    def rec(cfg, name, env, fn, nbytes):
        with Env(**env):
            ms = timeit(fn, a.reps)
            k = nv.last_launch()
        row = {"config": cfg, "variant": name, "env": env, "kernel": k, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
               "hbm_frac": round(nbytes / ms / 1e6 / hbm, 3)}
        print(json.dumps(row), flush=True)

    shapes = {"c2": (32 if a.small else 256, 3, 1024, 1024), "c4": (8 if a.small else 64, 3, 2160, 3840)}
    for cfg, shp in shapes.items():
        x = torch.rand(*shp, device="cuda") * 255
        y = torch.empty_like(x)
        n = x.numel()
        if "r2h" in a.what:
            fast = lambda: Fn.rect_to_hex(x, None, "bilinear", out_dtype=torch.float32, math="fast", out=y)
            exact = lambda: Fn.rect_to_hex(x, None, "bilinear", out_dtype=torch.float32, math="exact", out=y)
            rec(cfg, "r2h fast: TMA warp-specialised", {"HG_R2H_STREAM": 0}, fast, 8 * n)
            rec(cfg, "r2h exact f32: TMA warp-specialised", {"HG_R2H_STREAM": 0}, exact, 8 * n)
            for pf in (2, 4, 6, 8):
                rec(cfg, "r2h fast: stream v3 (batched loads)", {"HG_R2H_STREAM_PF": pf}, fast, 8 * n)
            for rows in (64,):
                for pf in (2, 4):
                    rec(cfg, "r2h fast: stream v2 (rolling prefetch)", {"HG_R2H_STREAM_V3": 0, "HG_R2H_STREAM_ROWS": rows, "HG_R2H_STREAM_PF": pf}, fast, 8 * n)
            for pf in (2, 4):
                rec(cfg, "r2h exact f32: stream v3", {"HG_R2H_STREAM_PF": pf}, exact, 8 * n)
            rec(cfg, "r2h exact f32: stream v2", {"HG_R2H_STREAM_V3": 0, "HG_R2H_STREAM_PF": 2}, exact, 8 * n)
            y64 = torch.empty(shp, device="cuda", dtype=torch.float64)
            e64 = lambda: Fn.rect_to_hex(x, None, "bilinear", out_dtype=torch.float64, math="exact", out=y64)
            rec(cfg, "r2h exact f64: direct gather", {"HG_R2H_STREAM": 0}, e64, 12 * n)
            for pf in (2, 4):
                rec(cfg, "r2h exact f64: stream v3", {"HG_R2H_STREAM_PF": pf}, e64, 12 * n)
            rec(cfg, "r2h exact f64: stream v2", {"HG_R2H_STREAM_V3": 0, "HG_R2H_STREAM_PF": 2}, e64, 12 * n)
            del y64
        if "h2r" in a.what:
            hf_ = lambda: Fn.hex_to_rect(x, None, "linear", out_dtype=torch.float32, math="fast", twin="np", out=y)
            he_ = lambda: Fn.hex_to_rect(x, None, "linear", out_dtype=torch.float32, math="exact", twin="np", out=y)
            for pf in (2, 4, 6):
                rec(cfg, "h2r fast: stream (batched loads)", {"HG_H2R_STREAM_PF": pf}, hf_, 8 * n)
            for warps in (8, 16):
                for R in (1, 2, 4):
                    rec(cfg, "h2r fast: TMA tiles", {"HG_H2R_STREAM": 0, "HG_HEXSRC_SHARE": R, "HG_HEXSRC_WARPS": warps}, hf_, 8 * n)
                for R in (8, 16, 32, 64):
                    rec(cfg, "h2r exact f32", {"HG_HEXSRC_SHARE": R, "HG_HEXSRC_WARPS": warps}, he_, 8 * n)
            for promo in (0, 1, 2, 3):
                rec(cfg, "h2r exact f32, L2 promotion", {"HG_HEXSRC_SHARE": 64, "HG_HEXSRC_L2PROMO": promo}, he_, 8 * n)
                rec(cfg, "h2r exact f32, L2 promotion", {"HG_HEXSRC_SHARE": 16, "HG_HEXSRC_L2PROMO": promo}, he_, 8 * n)
                rec(cfg, "h2r fast: TMA tiles, L2 promotion", {"HG_H2R_STREAM": 0, "HG_HEXSRC_L2PROMO": promo}, hf_, 8 * n)
        if "dist" in a.what:          # tile positions: contiguous ranges per CTA (0) against interleaved over the grid (1)
            hf_ = lambda: Fn.hex_to_rect(x, None, "linear", out_dtype=torch.float32, math="fast", twin="np", out=y)
            he_ = lambda: Fn.hex_to_rect(x, None, "linear", out_dtype=torch.float32, math="exact", twin="np", out=y)
            y64 = torch.empty(shp, device="cuda", dtype=torch.float64)
            he64 = lambda: Fn.hex_to_rect(x, None, "linear", math="exact", twin="np", out=y64)
            for dist in (0, 1):
                for R in (8, 16, 32, 64, 256):
                    rec(cfg, "h2r exact f32", {"HG_HEXSRC_DIST": dist, "HG_HEXSRC_SHARE": R}, he_, 8 * n)
                rec(cfg, "h2r exact f32, 16 warps", {"HG_HEXSRC_DIST": dist, "HG_HEXSRC_SHARE": 64, "HG_HEXSRC_WARPS": 16}, he_, 8 * n)
                rec(cfg, "h2r exact f32, column-major", {"HG_HEXSRC_DIST": dist, "HG_HEXSRC_SHARE": 64, "HG_HEXSRC_ORDER": 1}, he_, 8 * n)
                for R in (16, 64):
                    rec(cfg, "h2r exact f64", {"HG_HEXSRC_DIST": dist, "HG_HEXSRC_SHARE": R}, he64, 12 * n)
                for R in (1, 2, 4, 8):
                    rec(cfg, "h2r fast: TMA tiles", {"HG_H2R_STREAM": 0, "HG_HEXSRC_DIST": dist, "HG_HEXSRC_SHARE": R}, hf_, 8 * n)
            rec(cfg, "h2r exact f32 (shipped heuristic)", {}, he_, 8 * n)
            rec(cfg, "h2r exact f64 (shipped heuristic)", {}, he64, 12 * n)
            rec(cfg, "h2r fast: TMA tiles (shipped heuristic)", {"HG_H2R_STREAM": 0}, hf_, 8 * n)
            del y64
        if "h2r" in a.what:
            rec(cfg, "h2r fast (shipped heuristic)", {}, hf_, 8 * n)
            rec(cfg, "h2r exact f32 (shipped heuristic)", {}, he_, 8 * n)
        del x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
