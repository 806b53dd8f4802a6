#!/usr/bin/env python
"""Micro-benchmark of the hex-conv kernels on BASELINE config 3 (HexConv2d 64->64, radius 2, 256x256 lattice,
batch 128): forward / data-gradient / weight-gradient, direct stencil vs tcgen05, fp32 and bf16 activations.
CUDA events on the launching stream, 3 warm-ups, tensors (2.1 GB each in fp32) far larger than L2.

    python tools/bench_conv.py [--batch 128] [--reps 10] [--direct]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200"))
from HyGrid import _native as nv  # noqa: E402
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
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--cin", type=int, default=64)
    ap.add_argument("--cout", type=int, default=64)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--direct", action="store_true", help="also time the CUDA-core direct stencil (slow at 64x64)")
    ap.add_argument("--only", default="", help="comma list of ops to run (fwd,dgrad,wgrad); default all")
    ap.add_argument("--dtypes", default="", help="e.g. f32f32 to run a single activation dtype pair")
    a = ap.parse_args()
    N, Ci, Co, H = a.batch, a.cin, a.cout, a.hw
    dev = "cuda"
    torch.manual_seed(0)
    w = (torch.randn(Co, Ci, 1, 7, device=dev) * 0.05)
    bias = torch.randn(Co, device=dev)
    flops = 2.0 * 7 * Ci * Co * N * H * H
    res = {"config": f"HexConv2d {Ci}->{Co} r=2 s=1 pad=1, {N}x{Ci}x{H}x{H}", "flop_per_pass": flops, "rows": []}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm, tf = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops_sustained", 1400.0)
    pairs = ((torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32))
    if a.dtypes == "f32f32":
        pairs = pairs[:1]
    elif a.dtypes == "bf16bf16":
        pairs = pairs[1:2]
    for xdt, ydt in pairs:
        x = torch.randn(N, Ci, H, H, device=dev).to(xdt)
        gy = torch.randn(N, Co, H, H, device=dev).to(ydt)
        y = torch.empty(N, Co, H, H, device=dev, dtype=ydt)
        gx = torch.empty_like(x)
        gw = torch.zeros_like(w)
        gb = torch.zeros_like(bias)
        st = nv.stream_ptr(x.device)
        for algo in ([2, 1] if a.direct else [2]):
            d = nv.ConvDesc(N, Ci, Co, H, H, H, H, 2, 1, 1, 1, 1, 1, 0.0, nv.hg_dtype(xdt), nv.hg_dtype(ydt), algo, 0)
            ops = {"fwd": lambda: nv.call("hg_hexconv_fwd", C.byref(d), nv.ptr(x), nv.ptr(w), nv.ptr(bias), nv.ptr(y), st),
                   "dgrad": lambda: nv.call("hg_hexconv_dgrad", C.byref(d), nv.ptr(gy), nv.ptr(w), nv.ptr(gx), st)}
            if algo == 1 or not nv.query("hg_hexconv_umma_eligible", C.byref(d), 2):
                dw = nv.ConvDesc(N, Ci, Co, H, H, H, H, 2, 1, 1, 1, 1, 1, 0.0, nv.hg_dtype(xdt), nv.hg_dtype(ydt), 1, 0)
                if a.direct or algo == 2:
                    ops["wgrad(direct)"] = lambda: nv.call("hg_hexconv_wgrad", C.byref(dw), nv.ptr(x), nv.ptr(gy), nv.ptr(gw), nv.ptr(gb), st)
            else:
                ops["wgrad"] = lambda: nv.call("hg_hexconv_wgrad", C.byref(d), nv.ptr(x), nv.ptr(gy), nv.ptr(gw), nv.ptr(gb), st)
            if a.only:
                ops = {k: v for k, v in ops.items() if k.split("(")[0] in a.only.split(",")}
            for name, fn in ops.items():
                reps = a.reps if "direct" not in name and algo == 2 else max(1, a.reps // 5)
                ms = timeit(fn, reps)
                nbytes = x.numel() * x.element_size() + y.numel() * y.element_size()
                row = {"op": name, "algo": "tcgen05" if algo == 2 and "direct" not in name else "direct",
                       "x": str(xdt).split(".")[1], "y": str(ydt).split(".")[1], "ms": round(ms, 4),
                       "tflops": round(flops / ms / 1e9, 1), "tensor_frac": round(flops / ms / 1e9 / tf, 3),
                       "gbs": round(nbytes / ms / 1e6, 1), "hbm_frac": round(nbytes / ms / 1e6 / hbm, 3),
                       "hex_mpix_s": round(N * H * H / ms / 1e3, 1)}
                res["rows"].append(row)
                print(json.dumps(row), flush=True)
        del x, gy, y, gx
    print(json.dumps(res))


if __name__ == "__main__":
    main()
