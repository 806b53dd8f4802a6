"""Live cross-check of the oracle against the UNMODIFIED reference on seeded random shapes (build container only).

    python tests/golden/live_check.py        # exit code 0 = every comparison passed

Run as its own process (tests/test_oracle_live_reference.py does that): the reference package is called ``HyGrid`` like
ours, so it must not share ``sys.modules`` with the other tests.  Wider than the committed golden vectors: every case
here is a fresh shape / scale factor / dtype.  float64 results are compared with ``==``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as MG  # noqa: E402


def main():
    gnp, gt, hf, IMAGE, HEXIMAGE = MG.load_reference()
    from oracle import hygrid_oracle as O
    from oracle import hexframes_oracle as HO
    rng = np.random.default_rng(4711)
    checked = 0

    def eq(a, b, what):
        nonlocal checked
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
        assert np.array_equal(a, b, equal_nan=True), (what, float(np.nanmax(np.abs(a.astype(np.float64) - b))))
        checked += 1

    for _ in range(30):
        c, h, w = int(rng.integers(1, 5)), int(rng.integers(2, 70)), int(rng.integers(2, 90))
        h1, w1 = max(2, int(h * rng.uniform(0.4, 2.4))), max(2, int(w * rng.uniform(0.4, 2.4)))
        dt = ("u8", "f32", "f64")[int(rng.integers(0, 3))]
        img = MG.rand_img(rng, (c, h, w), dt)
        if c == 1:
            continue                                   # the reference squeezes single-band results; covered by the goldens
        for interp in ("nearest", "bilinear"):
            eq(O.rect_to_hex_resample(img, (h1, w1), interp), gnp.rect_to_hex_resample(img.copy(), (h1, w1), interp), f"R1 {interp} {img.shape}->{h1, w1} {dt}")
        eq(O.hex_to_rect_resample(img, (h1, w1), "linear", twin="np"), gnp.hex_to_rect_resample(img.copy(), (h1, w1), "linear"), f"R2 np {img.shape}")
        eq(O.hexresize(img, (h1, w1), "linear"), gnp.hexresize(img.copy(), (h1, w1), "linear"), f"R4 {img.shape}")
        for interp in ("nearest", "linear"):
            eq(O.hex_to_rect_resample(img, (h1, w1), interp, twin="torch"), gt.hex_to_square_resample(img.copy(), (h1, w1), interp), f"R2 torch {interp} {img.shape}")

    # R3 hex -> hex affine warp (numpy twin, float64 inverse map) and R5 doubled rasters
    for _ in range(10):
        c, h, w = int(rng.integers(2, 4)), int(rng.integers(6, 40)), int(rng.integers(6, 40))
        img = MG.rand_img(rng, (c, h, w), "f64")
        ang, sc = rng.uniform(-0.5, 0.5), rng.uniform(0.6, 1.6)
        Hm = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-3, 3)],
                       [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-3, 3)], [0, 0, 1.0]])
        eq(O.hex_warp(img, Hm, "linear", twin="np"), gnp.image_geometric_transformation(img.copy(), Hm, "linear"), f"R3 linear {img.shape}")
        hx = HEXIMAGE(data=img.copy(), even_odd_offset=bool(rng.integers(0, 2)))
        t1, _ = hx.GenerateType1Image()
        t2, _ = hx.GenerateType2Image()
        eq(O.hex_to_type1(img, int(hx.even_odd_offset)), t1, "R5 type1")
        eq(O.hex_to_type2(img, int(hx.even_odd_offset)), t2, "R5 type2")
        eq(O.type1_to_hex(t1), HEXIMAGE(data=t1.copy(), heximagetype=1).HexagonImage, "R5 decode1")
        eq(O.type2_to_hex(t2), HEXIMAGE(data=t2.copy(), heximagetype=2).HexagonImage, "R5 decode2")

    torch.manual_seed(4711)
    for _ in range(12):
        N, Cin, Cout = int(rng.integers(1, 3)), int(rng.integers(1, 6)), int(rng.integers(1, 6))
        H, W = int(rng.integers(6, 20)), int(rng.integers(6, 20))
        r, s_, d, pad, off = int(rng.integers(2, 4)), int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.integers(0, 3)), int(rng.integers(0, 2))
        try:
            m = hf.HexConv2d(Cin, Cout, off, r, stride=s_, padding=pad, dilation=d)
            x = torch.randn(N, Cin, H, W)
            yr = m(x)
        except Exception:
            continue                                   # shape too small for this kernel in the reference
        yo = HO.hexconv2d(x, m.kernel.detach(), m.bias.detach(), off, r, s_, pad, d, 1)
        assert yo.shape == yr.shape and float((yo - yr.detach()).abs().max()) <= 1e-5 * max(1.0, float(yr.abs().max())), "C1"
        checked += 1
    # even rows only: a padded height in [k_h, k_h + s) skips the odd conv and the reference returns the even-row result
    # alone (HexFrames.py:163-164): one output row
    for (H, W, r, s_, d, pad, off) in ((3, 9, 2, 1, 1, 0, 0), (3, 8, 2, 1, 1, 0, 1), (5, 12, 3, 1, 1, 0, 0), (1, 7, 2, 1, 1, 1, 0),
                                       (4, 11, 2, 2, 1, 0, 1), (5, 13, 2, 1, 2, 0, 0)):
        m = hf.HexConv2d(2, 3, off, r, stride=s_, padding=pad, dilation=d)
        x = torch.randn(2, 2, H, W)
        yr = m(x)
        yo = HO.hexconv2d(x, m.kernel.detach(), m.bias.detach(), off, r, s_, pad, d, 1)
        assert yr.shape[2] == 1, ("expected the even-rows-only case", yr.shape)
        assert yo.shape == yr.shape and float((yo - yr.detach()).abs().max()) <= 1e-5 * max(1.0, float(yr.abs().max())), "C1 even rows only"
        checked += 1
    for _ in range(12):
        B, C_, H, W = int(rng.integers(1, 3)), int(rng.integers(1, 4)), int(rng.integers(4, 24)), int(rng.integers(5, 30))
        method = ("max", "min", "average")[int(rng.integers(0, 3))]
        x = torch.randn(B, C_, H, W)
        yr = hf.HexPool2d(method, 2, 2)(x)
        yo = HO.hexpool2d(x, method, 2, 2)
        assert yo.shape == yr.shape and torch.allclose(yo, yr, atol=1e-6, equal_nan=True), "P1"
        checked += 1
    # P1 beyond 2x2: windows / strides / padding / ceil mode / NaN-aware reductions; error agreement where the window leaves the image
    for _ in range(40):
        B, C_, H, W = 1, int(rng.integers(1, 3)), int(rng.integers(6, 26)), int(rng.integers(8, 30))
        method = ("max", "min", "average")[int(rng.integers(0, 3))]
        kh, kw = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        sh, sw = int(rng.integers(1, 4)), int(rng.integers(2, 5))
        pad = int(rng.integers(0, 3))
        ceil, cip = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        x = torch.randn(B, C_, H, W)
        if rng.integers(0, 3) == 0:
            x[torch.rand_like(x) < 0.2] = float("nan")
        try:
            yr = hf.HexPool2d(method, (kh, kw), (sh, sw), padding=pad, ceil_mode=ceil, count_include_pad=cip)(x)
        except Exception:
            try:
                HO.hexpool2d(x, method, (kh, kw), (sh, sw), pad, "constant", 0, ceil, cip)
            except Exception:
                continue
            raise AssertionError(f"P1: the reference raises but the oracle does not ({method} k={kh, kw} s={sh, sw} pad={pad} ceil={ceil})")
        yo = HO.hexpool2d(x, method, (kh, kw), (sh, sw), pad, "constant", 0, ceil, cip)
        assert yo.shape == yr.shape and torch.allclose(yo, yr, atol=1e-6, equal_nan=True), f"P1 {method} k={kh, kw} s={sh, sw} pad={pad} ceil={ceil} cip={cip}"
        checked += 1
    # P2 adaptive / global pooling (the one undefined name of the reference is injected by make_golden.load_reference)
    for _ in range(16):
        H, W, n = int(rng.integers(6, 30)), int(rng.integers(8, 34)), int(rng.integers(1, 6))
        method = ("max", "min", "average")[int(rng.integers(0, 3))]
        x = torch.randn(2, 3, H, W)
        try:
            yr = hf.HexAdaptivePool2d(n, method)(x)
        except Exception:
            continue
        yo = HO.hexadaptivepool2d(x, n, method)
        assert yo.shape == yr.shape and torch.allclose(yo, yr, atol=1e-6, equal_nan=True), f"P2 adaptive {H, W}->{n} {method}"
        yg = hf.HexGlobalPool2d(method)(x)
        og = HO.hexglobalpool2d(x, method)
        assert og.shape == yg.shape and torch.allclose(og, yg, atol=1e-6), f"P2 global {method}"
        checked += 2
    # C1 with groups, C2 adaptive ("same") padding
    for _ in range(10):
        g = int(rng.integers(2, 4))
        Cin, Cout = g * int(rng.integers(1, 3)), g * int(rng.integers(1, 3))
        H, W, r, off = int(rng.integers(8, 18)), int(rng.integers(8, 18)), int(rng.integers(2, 4)), int(rng.integers(0, 2))
        m = hf.HexConv2d(Cin, Cout, off, r, stride=1, padding=r - 1, groups=g)
        x = torch.randn(1, Cin, H, W)
        yr = m(x)
        yo = HO.hexconv2d(x, m.kernel.detach(), m.bias.detach(), off, r, 1, r - 1, 1, g)
        assert yo.shape == yr.shape and float((yo - yr.detach()).abs().max()) <= 1e-5 * max(1.0, float(yr.abs().max())), "C1 groups"
        s_, d = int(rng.integers(1, 3)), int(rng.integers(1, 3))
        try:
            ma = hf.HexConv2dAdaptivePadding(Cin, Cout, off, r, stride=s_, dilation=d)
            ya = ma(x)
        except Exception:
            continue
        pl, pr, pt, pb = HO.adaptive_padding(H, W, r, s_, d)
        xa = torch.nn.functional.pad(x, (pl, pr, pt, pb))
        yo = HO.hexconv2d(xa, ma.kernel.detach(), ma.bias.detach(), off, r, s_, 0, d, 1)
        assert yo.shape == ya.shape and float((yo - ya.detach()).abs().max()) <= 1e-5 * max(1.0, float(ya.abs().max())), "C2 adaptive padding"
        checked += 2
    # retired layers ("codes in old versions.txt"), classes exec'd unmodified by make_retired_golden.load_class
    import make_retired_golden as MR
    PS, CT = MR.load_class()
    for _ in range(12):
        r, cout, H, W = int(rng.integers(2, 6)), int(rng.integers(1, 3)), int(rng.integers(1, 9)), int(rng.integers(1, 9))
        x = torch.randn(2, cout * r * r, H, W)
        eq(HO.hex_pixel_shuffle(x, r).numpy(), PS(r)(x).numpy(), f"pixel shuffle r={r} {H, W}")
    for _ in range(16):
        r, s_, off, H, W = int(rng.integers(2, 4)), int(rng.integers(1, 4)), int(rng.integers(0, 2)), int(rng.integers(3, 10)), int(rng.integers(3, 10))
        m = CT(3, 4, off, r, stride=s_, bias=True)
        x = torch.randn(1, 3, H, W)
        try:
            with torch.no_grad():
                yr = m(x)
        except Exception:
            try:
                HO.hex_conv_transpose2d(x, m.kernel.detach(), m.bias.detach(), off, r, s_)
            except ValueError:
                continue
            raise AssertionError(f"transposed conv: the reference raises but the oracle does not (r={r} s={s_} {H, W})")
        yo = HO.hex_conv_transpose2d(x, m.kernel.detach(), m.bias.detach(), off, r, s_)
        assert yo.shape == yr.shape and float((yo - yr).abs().max()) <= 1e-5 * max(1.0, float(yr.abs().max())), "transposed conv"
        checked += 1
    # heximpad (geometry_np.py:683-732) through cv2, with the one name the reference forgets to import injected
    import numbers
    import make_impad_golden as MI
    gnp.numbers = numbers
    for _ in range(60):
        img, kw = MI.random_case(rng)
        try:
            ref = gnp.heximpad(img.copy(), **kw)
        except Exception:
            try:
                O.heximpad(img, **kw)
            except Exception:
                continue
            raise AssertionError(f"heximpad: the reference raises but the oracle does not ({kw}, {img.shape}, {img.dtype})")
        eq(O.heximpad(img, **kw), ref, f"heximpad {kw} {img.shape} {img.dtype}")
    print(f"live reference check ok: {checked} comparisons")


if __name__ == "__main__":
    main()
