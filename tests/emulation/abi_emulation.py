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

NP = {nv.U8: np.uint8, nv.I16: np.int16, nv.I32: np.int32, nv.I64: np.int64, nv.F32: np.float32, nv.F64: np.float64, nv.U16: np.uint16}
CT = {np.uint8: C.c_uint8, np.int16: C.c_int16, np.int32: C.c_int32, np.int64: C.c_int64, np.float32: C.c_float, np.float64: C.c_double,
      np.uint16: C.c_uint16}


def view(ptr, n, dt):
    """numpy view of ``n`` elements of hg dtype ``dt`` at a raw pointer (bfloat16 as its uint16 bit pattern)."""
    addr = ptr.value if hasattr(ptr, "value") else ptr
    ct = C.c_uint16 if dt == nv.BF16 else CT[NP[dt]]
    if n == 0 or not addr:
        return np.zeros(0, np.uint16 if dt == nv.BF16 else NP[dt])       # empty tensors come with a null pointer
    return np.ctypeslib.as_array((ct * n).from_address(addr))


def bf16_to_f32(raw):
    return (raw.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def tensor(ptr, shape, dt):
    """float32 torch tensor holding the values of a raw buffer of hg dtype ``dt``."""
    n = int(np.prod(shape))
    raw = view(ptr, n, dt)
    vals = bf16_to_f32(raw) if dt == nv.BF16 else raw.astype(np.float32 if dt != nv.F64 else np.float64)
    return torch.from_numpy(np.ascontiguousarray(vals)).reshape(tuple(int(v) for v in shape))


def store(ptr, t, dt):
    flat = t.detach().reshape(-1)
    out = view(ptr, flat.numel(), dt)
    out[:] = f32_to_bf16(flat.float().numpy()) if dt == nv.BF16 else flat.numpy().astype(NP[dt])


last_launch = [""]


def conv_entry(name, a):
    """hg_hexconv_*: the oracle's closed form (and autograd through it) on the descriptor's geometry."""
    d = a[0]._obj
    eo = (d.parity - d.pad) % 2                # the oracle re-derives the parity from even_odd_offset + padding
    K = 3 * d.radius * d.radius - 3 * d.radius + 1
    wshape = (d.Cout, d.Cin // d.groups, 1, K)

    op = {"hg_hexconv_fwd": 0, "hg_hexconv_fwd_affine": 0, "hg_hexconv_dgrad": 1, "hg_hexconv_wgrad": 2}[name]
    tc = d.algo == 2 or (d.algo == 0 and umma_eligible(d, op))
    bf = (lambda t: t.bfloat16().float()) if tc else (lambda t: t)   # the tcgen05 kernels read activations and weights as bfloat16
    last_launch[0] = "hexconv_umma" if tc else "hexconv_direct"      # what hg_last_launch() would answer (kernel family)

    def run(x, w, b):                          # callers round the operands the kernel reads (never the autograd leaf)
        mode = ("constant", "reflect", "replicate", "circular")[d.pad_mode if d.pad else 0]      # header: frame filled in place
        return HO.hexconv2d(x, w, b, eo, d.radius, d.stride, d.pad, d.dilation, d.groups, mode, d.pad_value)
    if name in ("hg_hexconv_fwd", "hg_hexconv_fwd_affine"):
        if name == "hg_hexconv_fwd":
            _, x, w, b, y, _ = a
            scale = None
        else:
            _, x, w, scale, b, y, _ = a
        xt, wt = tensor(x, (d.N, d.Cin, d.H, d.W), d.x_dtype), tensor(w, wshape, nv.F32)
        if scale is not None and scale.value:       # folded into the weights on their way into shared memory (before the bf16 rounding)
            wt = wt * tensor(scale, (d.Cout,), nv.F32).view(-1, 1, 1, 1)
        out = run(bf(xt), bf(wt), None)
        if b is not None and b.value:
            out = out + tensor(b, (d.Cout,), nv.F32).view(1, -1, 1, 1)
        assert tuple(out.shape) == (d.N, d.Cout, d.Ho, d.Wo), (tuple(out.shape), d.Ho, d.Wo)
        if d.accumulate:                       # header: add to what y already holds (tcgen05 path only)
            out = out + tensor(y, (d.N, d.Cout, d.Ho, d.Wo), d.y_dtype)
        store(y, out.clamp_min(0) if d.relu else out, d.y_dtype)
    elif name == "hg_hexconv_dgrad":          # (the product's backward runs under no_grad: re-enable it for the oracle)
        _, gy, w, gx, _ = a
        xt = torch.zeros(d.N, d.Cin, d.H, d.W, requires_grad=True)
        with torch.enable_grad():
            run(xt, bf(tensor(w, wshape, nv.F32)), None).backward(bf(tensor(gy, (d.N, d.Cout, d.Ho, d.Wo), d.y_dtype)))
        store(gx, xt.grad + tensor(gx, (d.N, d.Cin, d.H, d.W), d.x_dtype) if d.accumulate else xt.grad, d.x_dtype)
    else:
        _, x, gy, gw, gb, _ = a
        wt = torch.zeros(wshape, requires_grad=True)
        bt = torch.zeros(d.Cout, requires_grad=True)
        with torch.enable_grad():
            run(bf(tensor(x, (d.N, d.Cin, d.H, d.W), d.x_dtype)), wt, bt).backward(bf(tensor(gy, (d.N, d.Cout, d.Ho, d.Wo), d.y_dtype)))
        view(gw, wt.numel(), nv.F32)[:] += wt.grad.reshape(-1).numpy()           # accumulates, like the kernel
        if gb is not None and gb.value:
            view(gb, d.Cout, nv.F32)[:] += bt.grad.numpy()


def bn_entry(name, a):
    """hg_bn_*: the formulas of include/hygrid_b200.h in float64 (float32 where the header says so)."""
    def f32(ptr, n):
        return view(ptr, n, nv.F32) if ptr is not None and getattr(ptr, "value", ptr) else None
    if name == "hg_bn_train_fwd":      # = zero the sums + hg_bn_stats + hg_bn_apply + BatchNorm2d's running-statistics update
        x, y, sums, gamma, beta, mean_out, rstd_out, rmean, rvar, nbt, momentum, N, Cn, HW, eps, relu, st = a
        view(sums, 2 * Cn, nv.F64)[:] = 0.0
        bn_entry("hg_bn_stats", (x, sums, N, Cn, HW, st))
        bn_entry("hg_bn_apply", (x, y, sums, None, None, gamma, beta, mean_out, None, rstd_out, N, Cn, HW, eps, relu, st))
        if f32(rmean, Cn) is not None:
            momentum = momentum.value if hasattr(momentum, "value") else float(momentum)
            if nbt is not None and getattr(nbt, "value", nbt):
                cnt = view(nbt, 1, nv.I64)
                cnt[0] += 1
            f = np.float32(momentum if momentum >= 0 else 1.0 / float(cnt[0]))
            sm = view(sums, 2 * Cn, nv.F64)
            n = N * HW
            mean = sm[0::2] / n
            var = np.maximum(sm[1::2] / n - mean * mean, 0.0)
            f32(rmean, Cn)[:] = f32(rmean, Cn) * (np.float32(1) - f) + f * mean.astype(np.float32)
            f32(rvar, Cn)[:] = f32(rvar, Cn) * (np.float32(1) - f) + (f * np.float32(n / (n - 1.0))) * var.astype(np.float32)
        return
    if name == "hg_bn_bwd":            # = zero the sums + hg_bn_bwd_reduce + hg_bn_bwd_apply (optionally accumulating dgamma / dbeta)
        x, dy, mean, rstd, gamma, beta, dsums, dx, dgamma, dbeta, acc, N, Cn, HW, relu, training, st = a
        view(dsums, 2 * Cn, nv.F64)[:] = 0.0
        bn_entry("hg_bn_bwd_reduce", (x, dy, mean, rstd, gamma, beta, dsums, N, Cn, HW, relu, st))
        old = [None if f32(p_, Cn) is None else f32(p_, Cn).copy() for p_ in (dgamma, dbeta)]
        bn_entry("hg_bn_bwd_apply", (x, dy, mean, rstd, gamma, beta, dsums, dx, dgamma, dbeta, N, Cn, HW, relu, training, st))
        if acc:
            for p_, o in zip((dgamma, dbeta), old):
                if o is not None:
                    f32(p_, Cn)[:] += o
        return
    if name == "hg_bn_stats":
        x, sums, N, Cn, HW, _ = a
        xs = view(x, N * Cn * HW, nv.F32).reshape(N, Cn, HW).astype(np.float64)
        out = view(sums, 2 * Cn, nv.F64)
        out[0::2] += xs.sum((0, 2))
        out[1::2] += (xs * xs).sum((0, 2))
        return
    if name == "hg_bn_apply":
        x, y, sums, mean_in, var_in, gamma, beta, mean_out, var_out, rstd_out, N, Cn, HW, eps, relu, _ = a
        eps = eps.value if hasattr(eps, "value") else float(eps)
        xs = view(x, N * Cn * HW, nv.F32).reshape(N, Cn, HW)
        if sums is not None and getattr(sums, "value", sums):
            sm = view(sums, 2 * Cn, nv.F64)
            n = N * HW
            mean = sm[0::2] / n
            var = np.maximum(sm[1::2] / n - mean * mean, 0.0)
        else:
            mean, var = f32(mean_in, Cn).astype(np.float64), f32(var_in, Cn).astype(np.float64)
        rstd = 1.0 / np.sqrt(var + eps)
        g, b = f32(gamma, Cn), f32(beta, Cn)
        scale = ((g if g is not None else 1.0) * rstd).astype(np.float32)
        shift = (-(mean.astype(np.float32)) * scale + (b if b is not None else 0.0)).astype(np.float32)
        out = xs * scale[None, :, None] + shift[None, :, None]
        view(y, N * Cn * HW, nv.F32)[:] = (np.maximum(out, 0) if relu else out).reshape(-1)
        for ptr, val in ((mean_out, mean), (var_out, var), (rstd_out, rstd)):
            if f32(ptr, Cn) is not None:
                f32(ptr, Cn)[:] = val.astype(np.float32)
        return
    if name == "hg_bn_bwd_reduce":
        x, dy, mean, rstd, gamma, beta, dsums, N, Cn, HW, relu, _ = a
        training = None
    else:
        x, dy, mean, rstd, gamma, beta, dsums, dx, dgamma, dbeta, N, Cn, HW, relu, training, _ = a
    xs = view(x, N * Cn * HW, nv.F32).reshape(N, Cn, HW).astype(np.float64)
    dz = view(dy, N * Cn * HW, nv.F32).reshape(N, Cn, HW).astype(np.float64)
    mu, rs = f32(mean, Cn).astype(np.float64)[None, :, None], f32(rstd, Cn).astype(np.float64)[None, :, None]
    g, b = f32(gamma, Cn), f32(beta, Cn)
    gam = (g.astype(np.float64) if g is not None else np.ones(Cn))[None, :, None]
    xhat = (xs - mu) * rs
    if relu:
        z = xhat * gam + (b.astype(np.float64) if b is not None else np.zeros(Cn))[None, :, None]
        dz = np.where(z > 0, dz, 0.0)
    if name == "hg_bn_bwd_reduce":
        out = view(dsums, 2 * Cn, nv.F64)
        out[0::2] += dz.sum((0, 2))
        out[1::2] += (dz * xhat).sum((0, 2))
        return
    ds = view(dsums, 2 * Cn, nv.F64)
    n = N * HW
    if training:
        res = gam * rs * (dz - ds[0::2][None, :, None] / n - xhat * ds[1::2][None, :, None] / n)
    else:
        res = gam * rs * dz
    view(dx, N * Cn * HW, nv.F32)[:] = res.reshape(-1).astype(np.float32)
    if f32(dgamma, Cn) is not None:
        f32(dgamma, Cn)[:] = ds[1::2].astype(np.float32)
    if f32(dbeta, Cn) is not None:
        f32(dbeta, Cn)[:] = ds[0::2].astype(np.float32)


real_call = nv.call
calls = []


def umma_eligible(d, op):
    """hg_hexconv_umma_eligible restated (csrc/hg_conv_umma.cu conv_umma_eligible, hg_conv_wgrad_umma.cu
    conv_wgrad_umma_eligible) without their shared-memory fit check, which asks the device."""
    if d.radius != 2 or d.stride != 1 or d.dilation != 1 or d.groups != 1:
        return False
    if op == 2:
        ok = 1 <= d.Cin <= 1024 and d.Cout % 8 == 0 and 8 <= d.Cout <= 1024        # input channels are rounded up to 16 in-kernel
        return ok and not (d.algo == 0 and (d.x_dtype != nv.BF16 or d.Cin * d.Cout < 1024))
    cred, nout = (d.Cin, d.Cout) if op == 0 else (d.Cout, d.Cin)
    if op == 0 and cred >= 1:
        cred = (cred + 15) // 16 * 16                 # the forward loader rounds the input channels up to 16
    if cred % 16 or not 16 <= cred <= 512 or nout % 16 or not 16 <= nout <= 256 or (op == 1 and d.relu):
        return False
    return not (d.algo == 0 and ((d.x_dtype if op == 0 else d.y_dtype) != nv.BF16 or cred * nout < 1024))


def real_query(name, dref, op):
    assert name == "hg_hexconv_umma_eligible"
    d = dref._obj
    forced = type(d).from_buffer_copy(d)
    return int(umma_eligible(forced, op))


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
    elif name in ("hg_hexconv_fwd", "hg_hexconv_fwd_affine", "hg_hexconv_dgrad", "hg_hexconv_wgrad"):
        conv_entry(name, a)
    elif name.startswith("hg_bn_"):
        bn_entry(name, a)
    elif name == "hg_pad2d_bwd":              # adjoint of the pad through autograd on F.pad
        gy, gx, planes, H, W, pl, pr, pt, pb, mode, dt, _ = a
        g = tensor(gy, (planes, H + pt + pb, W + pl + pr), dt)
        x = torch.zeros(1, planes, H, W, dtype=torch.float64, requires_grad=True)
        tmode = {0: "constant", 1: "reflect", 2: "replicate", 3: "circular"}[mode]
        with torch.enable_grad():
            torch.nn.functional.pad(x, (pl, pr, pt, pb), mode=tmode).backward(g[None].double())
        store(gx, x.grad[0], dt)
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
    elif name == "hg_broadcast_fill":
        dst, scalar, n, dt, _ = a
        store(dst, tensor(scalar, (1,), dt).expand(n), dt)
    elif name == "hg_split_bf16":             # hi = bf16(x), lo = bf16(x - hi)
        x, hi, lo, n, _ = a
        xs = torch.from_numpy(view(x, n, nv.F32).copy())
        h = xs.bfloat16()
        if hi is not None and getattr(hi, "value", hi):
            store(hi, h.float(), nv.BF16)
        if lo is not None and getattr(lo, "value", lo):
            store(lo, (xs - h.float()), nv.BF16)
    elif name == "hg_host_hex_to_type":       # host writer of the doubled rasters: same encoder, host pointers
        hexp, t, planes, H, W, off, sdt, ddt, rows_mul, _ = a
        s = view(hexp, planes * H * W, sdt).reshape(planes, H, W)
        r = np.asarray((O.hex_to_type1 if rows_mul == 1 else O.hex_to_type2)(s, off, NP[ddt]))
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
    from HyGrid import _norm
    _norm.bn_supported = lambda bn, x: (type(bn) is torch.nn.BatchNorm2d and isinstance(x, torch.Tensor) and x.dim() == 4
                                        and x.dtype == torch.float32 and x.numel() > 0
                                        and (bn.weight is None or bn.weight.dtype == torch.float32))   # = the product's rule minus is_cuda
    real_query = nv.query
    nv.query = lambda name, *a: 0 if name == "hg_hexconv_umma_eligible" else real_query(name, *a)   # no tcgen05 routing here


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
    import test_zz_hexconvmodule_variants as T7
    for order in T7.ORDERS:
        body(T7.test_every_order_with_batchnorm_and_relu)(order)
    print("ok HexConvModule: 6 orders x flags x train/eval x grad")
    params = [m.args[1] for m in T7.test_layer_types_padding_modes_and_flags.pytestmark if m.name == "parametrize"][0]
    for cfg in params:
        body(T7.test_layer_types_padding_modes_and_flags)(cfg)
    print(f"ok HexConvModule: {len(params)} layer / padding / flag variants")
    import test_zz_heximpad as T5
    body(T5.test_gpu_heximpad_returns_the_reference_arrays)(np.load(T5.GOLDEN)); print("ok heximpad")
    used = sorted(set(calls))
    assert {"hg_plane_gather", "hg_plane_scatter", "hg_host_hex2rect", "hg_hexwarp_linear", "hg_hexwarp_nearest", "hg_pad2d"} <= set(used), used
    print("emulated entry points:", ", ".join(used))


if __name__ == "__main__":
    main()
