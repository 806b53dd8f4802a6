"""Run the Python layer of the table-driven ops and of ``HyGrid.geometry`` on CPU tensors by emulating the handful of
C-ABI entry points they call (numpy for the gathers, the oracle for the resamplers and the hex conv).

    python tests/emulation/abi_emulation.py          # exit code 0 = every emulated GPU test body passed

TEST INFRASTRUCTURE: it checks the host-side glue -- argument order and types handed to the C ABI, shapes, dtype
handling, table construction, autograd wiring (this is how the missing accumulation of the non-injective selection
table of ``HexConvTranspose2d`` for even strides was found) -- NOT the kernels; those are only checked by the
``-m gpu`` tests on a B200.  It monkeypatches torch and the binding module process-wide, so it runs as its own process
(tests/test_zz_cpu_emulated_abi.py).  Each emulated entry point follows the contract in include/hygrid_b200.h.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from HyGrid import _native as nv, HexFrames as hf, functional as Fn, _hostapi   # noqa: E402
import HyGrid.geometry as _G, HyGrid.geometry_np as _GN, HyGrid.geometry_torch as _GT   # noqa: E402
from oracle import hexframes_oracle as HO, hygrid_oracle as O   # noqa: E402

NP = {nv.U8: np.uint8, nv.F32: np.float32, nv.F64: np.float64, nv.I64: np.int64, nv.U16: np.uint16}
CT = {np.uint8: C.c_uint8, np.float32: C.c_float, np.float64: C.c_double, np.int64: C.c_int64, np.uint16: C.c_uint16}


def view(ptr, n, dt):
    """numpy view of ``n`` elements of hg dtype ``dt`` at a raw pointer (bfloat16 as its uint16 bit pattern)."""
    addr = ptr.value if hasattr(ptr, "value") else ptr
    ct = C.c_uint16 if dt == nv.BF16 else CT[NP[dt]]
    return np.ctypeslib.as_array((ct * n).from_address(addr))


def bf16_to_f32(raw):
    return (raw.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


real_call = nv.call
calls = []


def emulated_call(name, *a):
    calls.append(name)
    if name == "hg_plane_gather":
        src, dst, tab, B, Cn, cells, bs, cs, sdt, ddt, _ = a
        t = view(tab, cells, nv.I64)
        s, d = view(src, B * bs, sdt), view(dst, B * Cn * cells, ddt)
        sf = bf16_to_f32(s) if sdt == nv.BF16 else s
        out = np.zeros((B, Cn, cells), np.float64)
        for b in range(B):
            for c in range(Cn):
                out[b, c] = np.where(t >= 0, sf[b * bs + c * cs + np.maximum(t, 0)], 0)
        d[:] = f32_to_bf16(out.reshape(-1).astype(np.float32)) if ddt == nv.BF16 else out.reshape(-1).astype(NP[ddt])
    elif name == "hg_plane_scatter":          # plain stores: a table that reads an element twice loses a contribution
        gd, gs, tab, B, Cn, cells, bs, cs, dt, _ = a
        t = view(tab, cells, nv.I64)
        g, o = view(gd, B * Cn * cells, dt).reshape(B, Cn, cells), view(gs, B * bs, dt)
        keep = t >= 0
        for b in range(B):
            for c in range(Cn):
                o[b * bs + c * cs + t[keep]] = g[b, c][keep]
    elif name == "hg_pad2d":
        x, y, planes, H, W, pl, pr, pt, pb, mode, value, dt, _ = a
        s = view(x, planes * H * W, dt).reshape(planes, H, W)
        np_mode = {0: "constant", 1: "reflect", 2: "edge", 3: "wrap", 4: "symmetric"}[mode]
        kw = {"constant_values": NP[dt](value)} if mode == 0 else {}
        view(y, planes * (H + pt + pb) * (W + pl + pr), dt)[:] = np.pad(s, ((0, 0), (pt, pb), (pl, pr)), mode=np_mode, **kw).reshape(-1)
    elif name == "hg_host_hex2rect":
        src, dst, xs, ys, c, h, w, h1, w1, sdt, ddt, interp, _, _ = a
        s = view(src, c * h * w, sdt).reshape(c, h, w)
        r = O.hexsrc_resample(s, view(xs, h1, nv.F64)[:, None], view(ys, w1, nv.F64)[None, :], interp)
        view(dst, c * h1 * w1, ddt)[:] = np.asarray(r).reshape(-1).astype(NP[ddt])
    elif name == "hg_host_rect2hex":          # the oracle's two stages on the caller's own coordinate tables
        src, dst, xs, ys, c, h, w, h1, w1, sdt, ddt, interp, _, _ = a
        s = view(src, c * h * w, sdt).reshape(c, h, w)
        i_n, i_f, j_n, j_f = O.rect2hex_index(h, w, view(xs, h1, nv.F64), view(ys, w1, nv.F64))
        def tap(ii, jj):
            ok = ((ii >= 0) & (ii < h))[:, None] & ((jj >= 0) & (jj < w))[None, :]
            return np.where(ok, s[:, np.clip(ii, 0, h - 1)][:, :, np.clip(jj, 0, w - 1)], 0)
        if interp == 0:
            r = tap(i_n, j_n)
        else:                                  # geometry_np.py:514-517
            p1, p2, p3, p4 = (tap(i_n, j_n).astype(np.float64), tap(i_n, j_n + 1).astype(np.float64),
                              tap(i_n + 1, j_n).astype(np.float64), tap(i_n + 1, j_n + 1).astype(np.float64))
            fi, fj = i_f[None, :, None], j_f[None, None, :]
            t1 = fi * p3 + (1 - fi) * p1
            t2 = fi * p4 + (1 - fi) * p2
            r = fj * t2 + (1 - fj) * t1
        view(dst, c * h1 * w1, ddt)[:] = np.asarray(r).reshape(-1).astype(NP[ddt])
    elif name in ("hg_hex_to_type1", "hg_hex_to_type2"):
        hexp, t, planes, H, W, off, sdt, ddt, _ = a
        s = view(hexp, planes * H * W, sdt).reshape(planes, H, W)
        fn = O.hex_to_type1 if name.endswith("1") else O.hex_to_type2
        r = np.asarray(fn(s, off, NP[ddt]))
        view(t, r.size, ddt)[:] = r.reshape(-1)
    elif name in ("hg_hexwarp_linear", "hg_hexwarp_nearest"):
        if name == "hg_hexwarp_linear":
            src, dst, cx, cy, f32, planes, h, w, h1, w1, sdt, ddt, _ = a
        else:
            src, dst, cx, cy, f32, planes, h, w, h1, w1, esz, _ = a
            sdt = ddt = {4: nv.F32, 8: nv.F64, 1: nv.U8}[esz]
        cdt = nv.F32 if f32 else nv.F64
        s = view(src, planes * h * w, sdt).reshape(planes, h, w)
        X, Y = view(cx, h1 * w1, cdt).reshape(h1, w1), view(cy, h1 * w1, cdt).reshape(h1, w1)
        r = O.hexsrc_resample(s, X, Y, 1 if name.endswith("linear") else 0, np.float32 if f32 else np.float64)
        view(dst, planes * h1 * w1, ddt)[:] = np.asarray(r).reshape(-1).astype(NP[ddt])
    else:
        return real_call(name, *a)            # shape queries run in the real library (no device needed)


def install():
    nv.call = emulated_call
    nv.require_cuda = lambda t, what="tensor": t
    nv.stream_ptr = lambda device=None: None
    torch.Tensor.cuda = lambda self, *a, **k: self.detach().clone()
    real_to = torch.Tensor.to

    def is_cuda_device(v):
        return (isinstance(v, torch.device) and v.type == "cuda") or (isinstance(v, str) and v.startswith("cuda"))

    def to(self, *a, **k):                    # .to(<cuda device>) stays on the CPU
        a = tuple("cpu" if is_cuda_device(v) else v for v in a)
        k = {key: ("cpu" if is_cuda_device(v) else v) for key, v in k.items()}
        return real_to(self, *a, **k)
    torch.Tensor.to = to
    torch.cuda.is_available = lambda: True
    torch.cuda.current_device = lambda: 0
    _hostapi.device_index = lambda device=None: 0

    def to_device(img, device=None):
        img = _hostapi.as_chw(img)
        if img.dtype not in _hostapi._INTERP_SRC:
            img = img.astype(np.float64)
        return torch.from_numpy(np.ascontiguousarray(img))
    for m in (_hostapi, _G, _GN, _GT):
        m.to_device = to_device
    # HexConvTranspose2d calls the hex conv through HexFrames.hexconv2d: the (differentiable) oracle stands in for the kernels
    hf.hexconv2d = lambda x, kernel, bias=None, even_odd_offset=0, radius=2, stride=1, padding=0, dilation=1, groups=1, **kw: \
        HO.hexconv2d(x.float(), kernel, bias, even_odd_offset, radius, stride, padding, dilation, groups)


def main():
    install()
    import test_zz_conv_transpose as T2
    import test_zz_hex_mosaic as T3
    import test_zz_numba_twin as T4
    import test_zz_pixel_shuffle as T1
    retired, numba = np.load(T1.GOLDEN), np.load(T4.GOLDEN)
    body = lambda f: getattr(f, "__wrapped__", f)   # noqa: E731
    body(T1.test_gpu_module_returns_the_reference_fixture)(retired); print("ok pixel shuffle: fixture")
    body(T1.test_gpu_ragged_shapes_dtypes_and_adjoint)(); print("ok pixel shuffle: ragged shapes, dtypes, adjoint")
    body(T2.test_gpu_module_matches_fixture_and_oracle)(retired); print("ok transposed conv: fixture, oracle, backward")
    body(T3.test_gpu_mosaic_equals_the_oracle_raster)(); print("ok hex mosaic")
    body(T4.test_gpu_module_returns_the_numba_fixture)(numba); print("ok numba twin: resampler and warp")
    body(T4.test_gpu_hexresize_follows_the_numpy_twin)(); print("ok numba twin: hexresize")
    import test_zz_numpy_api as T6
    golden = np.load(os.path.join(ROOT, "tests", "golden", "resample_golden.npz"))
    body(T6.test_geometry_np_against_the_reference_outputs)(golden); print("ok geometry_np")
    body(T6.test_geometry_torch_against_the_reference_outputs)(golden); print("ok geometry_torch")
    body(T6.test_error_behaviour_of_the_numpy_api)(); print("ok numpy api errors")
    body(T6.test_image_and_heximage_classes)(golden); print("ok IMAGE / HEXIMAGE")
    import test_zz_heximpad as T5
    body(T5.test_gpu_heximpad_returns_the_reference_arrays)(np.load(T5.GOLDEN)); print("ok heximpad")
    used = sorted(set(calls))
    assert {"hg_plane_gather", "hg_plane_scatter", "hg_host_hex2rect", "hg_hexwarp_linear", "hg_hexwarp_nearest", "hg_pad2d"} <= set(used), used
    print("emulated entry points:", ", ".join(used))


if __name__ == "__main__":
    main()
