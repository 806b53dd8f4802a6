#!/usr/bin/env python
"""Micro-benchmark of the hex-conv kernels on BASELINE config 3 (HexConv2d 64->64, radius 2, 256x256 lattice,
batch 128): forward / data-gradient / weight-gradient, direct stencil vs tcgen05, fp32 and bf16 activations.
CUDA events on the launching stream, 3 warm-ups, tensors (2.1 GB each in fp32) far larger than L2.

    python tools/bench_conv.py [--batch 128] [--reps 10] [--direct] [--cudnn]

--cudnn also times the route the REFERENCE takes for the same layer on the same GPU (HexFrames.py:96-169): expand the
7-tap kernel into a dense 3 x 5 window with structural zeros, materialise the doubled ("type1") image, run two strided
F.conv2d (cuDNN's sm_100 kernels) for the even and the odd output rows and interleave them -- float32 and under
autocast(bfloat16), forward and forward + backward -- and checks that it computes the same thing as hg_hexconv_fwd.
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


def cudnn_route(x, kernel, bias, pad=1, offset=0):
    """The reference's algorithm for radius 2, stride 1, dilation 1 (HexFrames.py:96-169), restated for timing."""
    F = torch.nn.functional
    Co, Ci = kernel.shape[:2]
    dense = kernel.new_zeros(Co, Ci, 3, 5)                      # (2r-1) x (4r-3) window, 7 of 15 taps used
    dense[:, :, 0, 1:4:2] = kernel[:, :, 0, 0:2]
    dense[:, :, 1, 0:5:2] = kernel[:, :, 0, 2:5]
    dense[:, :, 2, 1:4:2] = kernel[:, :, 0, 5:7]
    P = F.pad(x, (pad,) * 4)
    N, C_, Hp, Wp = P.shape
    o = (offset + pad) % 2
    two = P.repeat_interleave(2, dim=3)
    T = torch.zeros(N, C_, Hp, 2 * Wp + 1, device=x.device, dtype=torch.float32)     # the doubled image is float32 (:439-442)
    T[:, :, (1 - o)::2, 1:] = two[:, :, (1 - o)::2]             # rows with (i + o) odd: leading zero
    T[:, :, o::2, :-1] = two[:, :, o::2]                        # the others: trailing zero
    even = F.conv2d(T[:, :, :, 1:-1], dense, bias, stride=(2, 2))
    odd = F.conv2d(T[:, :, 1:, 2:], dense, bias, stride=(2, 2))
    y = torch.empty(N, Co, even.shape[2] + odd.shape[2], even.shape[3], device=x.device)   # float32 result (:157-160)
    y[:, :, ::2] = even[..., :y.shape[3]]
    y[:, :, 1::2] = odd[..., :y.shape[3]]
    return y


def bench_cudnn(a, res, flops, hbm, tf):
    N, Ci, Co, H = a.batch, a.cin, a.cout, a.hw
    dev = "cuda"
    torch.manual_seed(0)
    w = (torch.randn(Co, Ci, 1, 7, device=dev) * 0.05).requires_grad_()
    bias = torch.randn(Co, device=dev).requires_grad_()
    x = torch.randn(N, Ci, H, H, device=dev)
    # same op?  one small batch against the library's kernel
    with torch.no_grad():
        ours = hf.hexconv2d(x[:2], w.detach(), bias.detach(), 0, 2, 1, 1)
        theirs = cudnn_route(x[:2], w.detach(), bias.detach())
    agree = float((ours - theirs).abs().max()) / float(theirs.abs().max())
    print(json.dumps({"check": "cudnn route vs hg_hexconv_fwd (fp32, direct stencil)", "rel_err": agree}), flush=True)
    assert agree < 1e-3, agree
    xg = x.clone().requires_grad_()
    if not a.cudnn_only or True:
        # the CUDA-core direct stencil the float32 route replaced (round 1's float32 path), for the record
        hf.set_fp32_tensor_cores(False)
        md = hf.HexConv2d(Ci, Co, 0, 2, padding=1).cuda()

        def direct_fwd_bwd():
            md.kernel.grad = md.bias.grad = xg.grad = None
            y = md(xg)
            y.backward(torch.ones_like(y))
        ms_d = timeit(direct_fwd_bwd, 2)
        hf.set_fp32_tensor_cores(True)
        print(json.dumps({"op": "fwd+bwd", "mode": "f32", "ours_direct_stencil_ms": round(ms_d, 3)}), flush=True)
        del md
    for label, ac in (("f32", False), ("autocast-bf16", True)):
        def fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                return cudnn_route(x, w, bias)

        def fwd_bwd():
            w.grad = bias.grad = xg.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                y = cudnn_route(xg, w, bias)
            y.backward(torch.ones_like(y))
        m = hf.HexConv2d(Ci, Co, 0, 2, padding=1).cuda()

        def ours_fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                return m(x)

        def ours_fwd_bwd():
            m.kernel.grad = m.bias.grad = xg.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                y = m(xg)
            y.backward(torch.ones_like(y))
        for op, fn_ref, fn_ours, passes in (("fwd", fwd, ours_fwd, 1), ("fwd+bwd", fwd_bwd, ours_fwd_bwd, 3)):
            ms_ref = timeit(fn_ref, max(2, a.reps // 3))
            ms_ours = timeit(fn_ours, max(2, a.reps // 3))
            row = {"op": op, "mode": label, "cudnn_route_ms": round(ms_ref, 3), "ours_module_ms": round(ms_ours, 3),
                   "ratio_cudnn_over_ours": round(ms_ref / ms_ours, 2),
                   "cudnn_route_tflops": round(passes * flops / ms_ref / 1e9, 1), "ours_tflops": round(passes * flops / ms_ours / 1e9, 1),
                   "note": "reference route = F.pad + type1 materialisation + 2 x F.conv2d (cuDNN) + interleave; ours = HexConv2d module"
                           + (" (tcgen05 kernels)" if ac else " (tcgen05, three bfloat16 passes over split operands: float32-class accuracy)")}
            res["rows"].append(row)
            print(json.dumps(row), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--cin", type=int, default=64)
    ap.add_argument("--cout", type=int, default=64)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--w", type=int, default=0, help="lattice width when it differs from --hw (the pooled layers of C5: 64 x 63, 32 x 31)")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--direct", action="store_true", help="also time the CUDA-core direct stencil (slow at 64x64)")
    ap.add_argument("--only", default="", help="comma list of ops to run (fwd,dgrad,wgrad); default all")
    ap.add_argument("--dtypes", default="", help="e.g. f32f32 to run a single activation dtype pair")
    ap.add_argument("--cudnn", action="store_true", help="also time the reference's cuDNN route (HexFrames.py:96-169) on this GPU")
    ap.add_argument("--cudnn-only", action="store_true")
    a = ap.parse_args()
    N, Ci, Co, H = a.batch, a.cin, a.cout, a.hw
    W = a.w or H
    dev = "cuda"
    torch.manual_seed(0)
    w = (torch.randn(Co, Ci, 1, 7, device=dev) * 0.05)
    bias = torch.randn(Co, device=dev)
    flops = 2.0 * 7 * Ci * Co * N * H * W
    res = {"config": f"HexConv2d {Ci}->{Co} r=2 s=1 pad=1, {N}x{Ci}x{H}x{W}", "flop_per_pass": flops, "rows": []}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm, tf = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops_sustained", 1400.0)
    if a.cudnn or a.cudnn_only:
        bench_cudnn(a, res, flops, hbm, tf)
        torch.cuda.empty_cache()
    if a.cudnn_only:
        print(json.dumps(res))
        return
    pairs = ((torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32))
    if a.dtypes == "f32f32":
        pairs = pairs[:1]
    elif a.dtypes == "bf16bf16":
        pairs = pairs[1:2]
    for xdt, ydt in pairs:
        x = torch.randn(N, Ci, H, W, device=dev).to(xdt)
        gy = torch.randn(N, Co, H, W, device=dev).to(ydt)
        y = torch.empty(N, Co, H, W, device=dev, dtype=ydt)
        gx = torch.empty_like(x)
        gw = torch.zeros_like(w)
        gb = torch.zeros_like(bias)
        st = nv.stream_ptr(x.device)
        for algo in ([2, 1] if a.direct else [2]):
            d = nv.ConvDesc(N, Ci, Co, H, W, H, W, 2, 1, 1, 1, 1, 1, 0.0, nv.hg_dtype(xdt), nv.hg_dtype(ydt), algo, 0)
            ops = {"fwd": lambda: nv.call("hg_hexconv_fwd", C.byref(d), nv.ptr(x), nv.ptr(w), nv.ptr(bias), nv.ptr(y), st),
                   "dgrad": lambda: nv.call("hg_hexconv_dgrad", C.byref(d), nv.ptr(gy), nv.ptr(w), nv.ptr(gx), st)}
            if algo == 1 or not nv.query("hg_hexconv_umma_eligible", C.byref(d), 2):
                dw = nv.ConvDesc(N, Ci, Co, H, W, H, W, 2, 1, 1, 1, 1, 1, 0.0, nv.hg_dtype(xdt), nv.hg_dtype(ydt), 1, 0)
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
                       "hex_mpix_s": round(N * H * W / ms / 1e3, 1)}
                res["rows"].append(row)
                print(json.dumps(row), flush=True)
        del x, gy, y, gx
    print(json.dumps(res))


if __name__ == "__main__":
    main()
