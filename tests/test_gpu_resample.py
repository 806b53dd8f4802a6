"""GPU parity tests of the rect<->hex resampling path.  Every test calls the sm_100a kernels through the
C ABI (ctypes -> libhygrid_b200.so) and compares with (a) the golden fixtures generated from the
reference (tests/golden/*.npz), (b) the CPU oracle on seeded inputs, (c) size-independent properties at
the BASELINE.json sizes.

Tolerances: integer tables / nearest / layout results bit-exact; float64 results of the EXACT math mode
bit-exact (== the reference's float64 outputs); float32 / FAST results within 1e-5 * max|x|.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import hygrid_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Fn():
    from HyGrid import functional
    return functional


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def same(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b, equal_nan=True), float(np.nanmax(np.abs(a.astype(np.float64) - b)))


def close(a, b, scale):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    assert a.shape == np.asarray(b).shape
    err = float(np.max(np.abs(a.astype(np.float64) - np.asarray(b, dtype=np.float64)))) if a.size else 0.0
    assert err <= 1e-5 * max(scale, 1e-30), (err, scale)


def _dsize(a):
    return None if a[0] < 0 else (int(a[0]), int(a[1]))


def index_image(h, w):
    ii, jj = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    return np.stack([ii + 1, jj + 1], 0).astype(np.float64)


# ------------------------------------------------------------------------------------------------
# R1 rect -> hex
# ------------------------------------------------------------------------------------------------
def test_r1_golden_bit_exact(Fn, resample_golden):
    G = resample_golden
    assert int(G["r1_count"]) >= 12
    for n in range(int(G["r1_count"])):
        img, ds, interp = G[f"r1_{n}_img"], _dsize(G[f"r1_{n}_dsize"]), str(G[f"r1_{n}_interp"])
        out = Fn.rect_to_hex(cu(img), ds, interp)                 # exact math, reference result dtype
        same(out.squeeze(), G[f"r1_{n}_out"])
        if interp == "bilinear":
            scale = float(np.abs(img).max())
            close(Fn.rect_to_hex(cu(img), ds, interp, out_dtype=torch.float32).squeeze(), G[f"r1_{n}_out"], scale)
            if img.dtype != np.float64:
                close(Fn.rect_to_hex(cu(img), ds, interp, out_dtype=torch.float32, math="fast").squeeze(), G[f"r1_{n}_out"], scale)


def test_r1_index_tables(Fn, resample_golden):
    G = resample_golden
    for n in range(int(G["r1idx_count"])):
        h, w, h1, w1 = (int(v) for v in G[f"r1idx_{n}_hw"])
        i_n, i_f, j_n, j_f = Fn.rect2hex_index(h, w, h1, w1)
        xs, ys = O.rect2hex_coords(h, w, h1, w1)
        oi, oif, oj, ojf = O.rect2hex_index(h, w, xs, ys)
        same(i_n, oi.astype(np.int32)); same(j_n, oj.astype(np.int32))
        same(i_f, oif); same(j_f, ojf)
        # the reference's own gather table, revealed by pushing an index image through 'nearest'
        out = Fn.rect_to_hex(cu(index_image(h, w)), (h1, w1), "nearest").cpu().numpy().astype(np.int32)
        assert np.array_equal(out[0].max(axis=1), G[f"r1idx_{n}_i"])
        assert np.array_equal(out[1].max(axis=0), G[f"r1idx_{n}_j"])


@pytest.mark.parametrize("dtype", [np.uint8, np.float32, np.float64])
@pytest.mark.parametrize("shape,dsize", [((3, 61, 47), None), ((3, 61, 47), (30, 23)), ((2, 33, 130), (70, 257)),
                                         ((1, 200, 300), (129, 500)), ((4, 5, 7), (1, 1)), ((3, 2, 2), (5, 5))])
def test_r1_vs_oracle(Fn, dtype, shape, dsize):
    rng = np.random.default_rng(7)
    img = (rng.random(shape) * 255).astype(dtype)
    for interp in ("nearest", "bilinear"):
        ref = O.rect_to_hex_resample(img, dsize, interp).reshape((shape[0],) + (dsize or shape[1:]))
        out = Fn.rect_to_hex(cu(img), dsize, interp)
        same(out, ref)
        if interp == "bilinear" and dtype != np.float64:
            close(Fn.rect_to_hex(cu(img), dsize, interp, out_dtype=torch.float32, math="fast"), ref, 255.0)


def test_r1_batched_equals_per_image(Fn):
    x = torch.rand(5, 3, 70, 90, device="cuda") * 255
    y = Fn.rect_to_hex(x, (64, 100), "bilinear", out_dtype=torch.float32, math="fast")
    for n in range(5):
        assert torch.equal(y[n], Fn.rect_to_hex(x[n], (64, 100), "bilinear", out_dtype=torch.float32, math="fast"))
    ref = O.rect_to_hex_resample(x[3].cpu().numpy(), (64, 100), "bilinear")
    close(y[3], ref, 255.0)


def test_r1_empty_and_noncontiguous(Fn):
    assert Fn.rect_to_hex(torch.zeros(0, 3, 8, 8, device="cuda"), None, "bilinear").shape == (0, 3, 8, 8)
    x = torch.rand(3, 40, 50, device="cuda", dtype=torch.float64)
    xt = x.transpose(1, 2)                                   # non-contiguous (3, 50, 40)
    same(Fn.rect_to_hex(xt, None, "bilinear"), O.rect_to_hex_resample(xt.cpu().numpy(), None, "bilinear"))
    with pytest.raises(KeyError):
        Fn.rect_to_hex(x, None, "linear")                    # the reference's method_dict KeyError


def test_r1_full_size_properties(Fn):
    """BASELINE config 2 geometry (3x1024x1024 f32 -> 1024x1024), reduced batch: linearity, constants,
    and the exact gather table through the index-image trick."""
    torch.manual_seed(0)
    a = torch.rand(4, 3, 1024, 1024, device="cuda") * 255
    b = torch.rand(4, 3, 1024, 1024, device="cuda") * 255
    f = lambda t: Fn.rect_to_hex(t, (1024, 1024), "bilinear", out_dtype=torch.float32, math="fast")
    ya, yb, yab = f(a), f(b), f(0.25 * a + 0.5 * b)
    assert float((yab - (0.25 * ya + 0.5 * yb)).abs().max()) <= 1e-5 * 255 * 2
    ones = f(torch.ones(1, 1, 1024, 1024, device="cuda"))
    assert float((ones[..., 1:-1, 1:-1] - 1).abs().max()) <= 1e-5      # interior weights sum to 1
    assert float(ones[..., :, 0].abs().max()) == 0 and float(ones[..., :, -1].abs().max()) == 0   # zero-filled columns
    # one image checked against the oracle in full
    ref = O.rect_to_hex_resample(a[1].cpu().numpy(), (1024, 1024), "bilinear")
    close(ya[1], ref, 255.0)
    same(Fn.rect_to_hex(a[1].double(), (1024, 1024), "bilinear"), O.rect_to_hex_resample(a[1].double().cpu().numpy(), (1024, 1024), "bilinear"))
    idx = Fn.rect_to_hex(cu(index_image(1024, 1024)), (1024, 1024), "nearest").cpu().numpy()
    xs, ys = O.rect2hex_coords(1024, 1024, 1024, 1024)
    i_n, _, j_n, _ = O.rect2hex_index(1024, 1024, xs, ys)
    ok = ((i_n >= 0) & (i_n < 1024))[:, None] & ((j_n >= 0) & (j_n < 1024))[None, :]
    assert np.array_equal(idx[0], np.where(ok, i_n[:, None] + 1, 0))
    assert np.array_equal(idx[1], np.where(ok, j_n[None, :] + 1, 0))


# ------------------------------------------------------------------------------------------------
# R2 / R4 hex -> rect, hexresize
# ------------------------------------------------------------------------------------------------
def test_config1_round_trip(Fn, resample_golden):
    G = resample_golden
    import hashlib
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    img = np.random.default_rng(0).integers(0, 256, (3, 512, 512), dtype=np.uint8)
    hexd = Fn.rect_to_hex(cu(img), (256, 256), "nearest")
    assert sha(hexd.cpu().numpy()) == str(G["c1_hex_sha"])
    back = Fn.hex_to_rect(hexd, (512, 512), "linear", twin="np")
    assert back.dtype == torch.float64
    assert sha(back.cpu().numpy()) == str(G["c1_back_sha"])
    same(back[:, ::37, ::41].contiguous(), G["c1_back_probe"])


def test_r2_r4_golden(Fn, resample_golden):
    G = resample_golden
    assert int(G["r2_count"]) >= 4
    for n in range(int(G["r2_count"])):
        img, ds = G[f"r2_{n}_img"], _dsize(G[f"r2_{n}_dsize"])
        x = cu(img)
        same(Fn.hex_to_rect(x, ds, "linear", twin="np").squeeze(), G[f"r2_{n}_np_linear"])
        same(Fn.hex_to_rect(x, ds, "linear", twin="torch").squeeze(), G[f"r2_{n}_torch_linear"])
        same(Fn.hex_to_rect(x, ds, "nearest", twin="torch").squeeze(), G[f"r2_{n}_torch_nearest"])
        same(Fn.hex_resize(x, ds if ds else img.shape[-2:], "linear").squeeze(), G[f"r2_{n}_resize_linear"])
        scale = float(np.abs(img).max())
        if img.dtype != np.float64:
            close(Fn.hex_to_rect(x, ds, "linear", out_dtype=torch.float32, math="fast", twin="np").squeeze(), G[f"r2_{n}_np_linear"], scale)


def test_r2_index_tables(Fn, resample_golden):
    G = resample_golden
    import hashlib
    for n in range(int(G["r2idx_count"])):
        h, w, h1, w1 = (int(v) for v in G[f"r2idx_{n}_hw"])
        out = Fn.hex_to_rect(cu(index_image(h, w)), (h1, w1), "nearest", twin="torch").cpu().numpy().astype(np.int32)
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == str(G[f"r2idx_{n}_sha"])
        for twin in ("np", "torch"):
            xs, ys = O.hex2rect_coords(h, w, h1, w1, twin)
            r = O.hexsrc_index(h, w, xs[:, None], ys[None, :])
            i_n, j_n, tri, off = Fn.hexsrc_index(h, w, cu(xs), cu(ys))
            same(i_n, r["i_n"].astype(np.int32)); same(j_n, r["j_n"].astype(np.int32))
            assert np.array_equal((tri.cpu().numpy() & 1).astype(bool), r["flag"])

            def lin(i, j):
                ok = (i >= 0) & (i < h) & (j >= 0) & (j < w)
                return np.where(ok, i * w + j, -1).astype(np.int32)
            f = r["flag"]
            exp = np.stack([lin(r["i_1"], r["j_1"]),
                            np.where(f, lin(r["i_2"], r["j_2"]), lin(r["i_3"], r["j_3"])),
                            lin(r["i_4"], r["j_4"])], 0)
            same(off, exp)


@pytest.mark.parametrize("dtype", [np.uint8, np.float32, np.float64])
@pytest.mark.parametrize("shape,dsize", [((3, 31, 29), (62, 58)), ((2, 40, 130), (33, 257)), ((1, 100, 120), None),
                                         ((3, 4, 4), (9, 1)), ((9, 16, 16), (32, 32))])
def test_r2_r4_vs_oracle(Fn, dtype, shape, dsize):
    rng = np.random.default_rng(11)
    img = (rng.random(shape) * 255).astype(dtype)
    full = lambda a: np.asarray(a).reshape((shape[0],) + tuple(dsize or shape[1:]))
    for twin in ("np", "torch"):
        same(Fn.hex_to_rect(cu(img), dsize, "linear", twin=twin), full(O.hex_to_rect_resample(img, dsize, "linear", twin=twin)))
        same(Fn.hex_to_rect(cu(img), dsize, "nearest", twin=twin), full(O.hex_to_rect_resample(img, dsize, "nearest", twin=twin)))
    ds = dsize or shape[1:]
    same(Fn.hex_resize(cu(img), ds, "linear"), full(O.hexresize(img, ds, "linear")))
    same(Fn.hex_resize(cu(img), ds, "nearest"), full(O.hexresize(img, ds, "nearest")))
    if dtype != np.float64:
        close(Fn.hex_to_rect(cu(img), dsize, "linear", out_dtype=torch.float32, math="fast", twin="np"),
              full(O.hex_to_rect_resample(img, dsize, "linear", twin="np")), 255.0)
    with pytest.raises(NotImplementedError):
        Fn.hex_to_rect(cu(img), dsize, "bilinear")


@pytest.mark.parametrize("shape,dsize", [((5, 64, 256), None), ((4, 96, 132), (200, 300)), ((1, 100, 120), (100, 120)),
                                         ((7, 33, 16), (70, 40)), ((3, 50, 260), (47, 250)), ((6, 40, 64), (17, 300))])
def test_r2_r4_tma_tiles_vs_oracle(Fn, shape, dsize):
    """float32 HG_MATH_FAST hex-source resampling on 16-byte aligned rows runs the TMA-staged tile kernel
    (hexsrc_linear_tma): plane groups of 3 with a ragged last group, tiles cut by the image border, up- and
    down-sampling; same lattice points as the oracle, float32 weights (<= 1e-5 * max|x|)."""
    rng = np.random.default_rng(13)
    img = (rng.random(shape) * 255).astype(np.float32)
    ds = dsize or shape[1:]
    full = lambda a: np.asarray(a).reshape((shape[0],) + tuple(ds))
    close(Fn.hex_to_rect(cu(img), dsize, "linear", out_dtype=torch.float32, math="fast", twin="np"),
          full(O.hex_to_rect_resample(img, dsize, "linear", twin="np")), 255.0)
    close(Fn.hex_resize(cu(img), ds, "linear", out_dtype=torch.float32, math="fast"), full(O.hexresize(img, ds, "linear")), 255.0)
    # HG_MATH_EXACT on the same tiles: float64 result bit-identical to the reference, float32 result = its rounding
    for twin in ("np", "torch"):
        ref = full(O.hex_to_rect_resample(img, dsize, "linear", twin=twin))
        same(Fn.hex_to_rect(cu(img), dsize, "linear", twin=twin), ref)
        same(Fn.hex_to_rect(cu(img), dsize, "linear", out_dtype=torch.float32, twin=twin), ref.astype(np.float32))
    same(Fn.hex_resize(cu(img), ds, "linear"), full(O.hexresize(img, ds, "linear")))
    # a constant image stays constant wherever all three lattice points are inside (weights sum to one)
    ones = Fn.hex_to_rect(torch.ones(shape, device="cuda"), dsize, "linear", out_dtype=torch.float32, math="fast", twin="np")
    ref1 = full(O.hex_to_rect_resample(np.ones(shape, np.float32), dsize, "linear", twin="np"))
    close(ones, ref1, 1.0)


def test_r2_full_size_properties(Fn):
    """4K geometry (config 4), one image: linearity + partition of unity of the barycentric weights."""
    torch.manual_seed(1)
    a = torch.rand(1, 3, 2160, 3840, device="cuda")
    b = torch.rand(1, 3, 2160, 3840, device="cuda")
    f = lambda t: Fn.hex_to_rect(t, (2160, 3840), "linear", out_dtype=torch.float32, math="fast", twin="np")
    assert float((f(0.5 * a - 0.25 * b) - (0.5 * f(a) - 0.25 * f(b))).abs().max()) <= 2e-5
    ones = f(torch.ones(1, 1, 2160, 3840, device="cuda"))
    assert float((ones[..., 2:-2, 2:-2] - 1).abs().max()) <= 1e-5
    hexed = Fn.rect_to_hex(a, None, "bilinear", out_dtype=torch.float32, math="fast")
    assert hexed.shape == a.shape and bool(torch.isfinite(hexed).all())


# ------------------------------------------------------------------------------------------------
# R3 warp
# ------------------------------------------------------------------------------------------------
def test_r3_golden(Fn, resample_golden):
    G = resample_golden
    assert int(G["r3_count"]) >= 3
    for n in range(int(G["r3_count"])):
        img, H = G[f"r3_{n}_img"], G[f"r3_{n}_H"]
        x = cu(img)
        same(Fn.hex_warp(x, H, "linear", twin="np").squeeze(), G[f"r3_{n}_np_linear"])
        same(Fn.hex_warp(x, H, "nearest", twin="torch").squeeze(), G[f"r3_{n}_torch_nearest"])
        out = Fn.hex_warp(x, H, "linear", twin="torch").squeeze()
        ref = G[f"r3_{n}_torch_linear"]
        assert out.shape == ref.shape
        same(out.to(torch.from_numpy(ref).dtype), ref)


def test_r3_identity_and_affine_kernel(Fn):
    rng = np.random.default_rng(3)
    img = rng.random((3, 19, 23))
    same(Fn.hex_warp(cu(img), np.eye(3), "linear", twin="np"), img)
    th = 0.3
    H = np.array([[np.cos(th), -np.sin(th), 1.5], [np.sin(th), np.cos(th), -2.0], [0, 0, 1.0]])
    ref = O.hex_warp(img, H, "linear", twin="np")
    same(Fn.hex_warp(cu(img), H, "linear", twin="np"), ref)
    # in-kernel inverse map: same lattice, coordinates may differ in the last ulp -> toleranced
    out = Fn.hex_warp_affine(cu(img), H, "linear", coord_f32=False)
    d = np.abs(out.cpu().numpy() - ref)
    assert np.quantile(d, 0.99) <= 1e-9 and out.shape == ref.shape
    f32 = O.hex_warp(img.astype(np.float32), H, "linear", twin="torch")
    o32 = Fn.hex_warp(cu(img.astype(np.float32)), H, "linear", twin="torch")
    same(o32, f32)


# ------------------------------------------------------------------------------------------------
# lattice index helpers and doubled rasters
# ------------------------------------------------------------------------------------------------
def test_axial_offset(Fn):
    i = torch.arange(-7, 60, device="cuda", dtype=torch.int32)[:, None]
    j = torch.arange(-9, 50, device="cuda", dtype=torch.int32)[None, :]
    ax = Fn.offset_to_axial(i, j)
    same(ax, O.offset_to_axial(i.cpu().numpy(), j.cpu().numpy()).astype(np.int32))
    same(Fn.axial_to_offset(i, ax), np.broadcast_to(j.cpu().numpy(), ax.shape).astype(np.int32))
    same(Fn.axial_to_offset(i, j), O.axial_to_offset(i.cpu().numpy(), j.cpu().numpy()).astype(np.int32))


def test_r5_doubled_rasters(Fn, resample_golden):
    G = resample_golden
    assert int(G["r5_count"]) >= 2
    for n in range(int(G["r5_count"])):
        img, off = G[f"r5_{n}_img"], int(G[f"r5_{n}_off"])
        x = cu(img)
        same(Fn.hex_to_type1(x, off, out_dtype=torch.float64), G[f"r5_{n}_t1"])
        same(Fn.hex_to_type2(x, off, out_dtype=torch.float64), G[f"r5_{n}_t2"])
        same(Fn.type1_to_hex(cu(G[f"r5_{n}_t1"])), G[f"r5_{n}_dec1"])
        same(Fn.type2_to_hex(cu(G[f"r5_{n}_t2"])), G[f"r5_{n}_dec2"])
        same(Fn.hex_to_type1(x.float()[None], off), G[f"r5_{n}_tt1"])
        same(Fn.hex_to_type2(x.float()[None], off), G[f"r5_{n}_tt2"])


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32, torch.float64, torch.bfloat16, torch.int16])
def test_r5_round_trip_large(Fn, dtype):
    x = (torch.rand(3, 5, 301, 517, device="cuda") * 200).to(dtype)
    for off in (0, 1):
        t1 = Fn.hex_to_type1(x, off)
        assert t1.shape == (3, 5, 301, 1035) and t1.dtype == dtype
        assert torch.equal(Fn.type1_to_hex(t1), x)
        t2 = Fn.hex_to_type2(x, off)
        assert torch.equal(Fn.type2_to_hex(t2), x)
        ref = O.hex_to_type1(x[0, 0].float().cpu().numpy(), off, np.float32)
        assert np.array_equal(t1[0, 0].float().cpu().numpy(), ref)


# ------------------------------------------------------------------------------------------------
# host-buffer entry points (numpy in / numpy out through the pinned, chunked pipeline)
# ------------------------------------------------------------------------------------------------
def test_host_entry_points(Fn):
    from HyGrid import _native as nv
    rng = np.random.default_rng(5)
    src = (rng.random((40, 96, 130)) * 255).astype(np.float32)
    h, w, h1, w1 = 96, 130, 80, 200
    for kind, name, coords in ((0, "hg_host_rect2hex", O.rect2hex_coords(h, w, h1, w1)),
                               (1, "hg_host_hex2rect", O.hex2rect_coords(h, w, h1, w1))):
        xs, ys = (np.ascontiguousarray(c) for c in coords)
        for pinned in (False, True):
            s = torch.from_numpy(src)
            d = torch.empty((40, h1, w1), dtype=torch.float64)
            if pinned:
                s, d = s.pin_memory(), d.pin_memory()
            nv.call(name, C.c_void_p(s.data_ptr()), C.c_void_p(d.data_ptr()), C.c_void_p(xs.ctypes.data),
                    C.c_void_p(ys.ctypes.data), 40, h, w, h1, w1, nv.F32, nv.F64, 1, nv.MATH_EXACT, 0)
            ref = (O.rect_to_hex_resample(src, (h1, w1), "bilinear") if kind == 0
                   else O.hex_to_rect_resample(src, (h1, w1), "linear"))
            same(d.numpy(), ref)
    nv.lib().hg_host_release()


# ------------------------------------------------------------------------------------------------
# seeded shape fuzz over every resampling kernel family (TMA tiles, direct gathers, exact / fast math)
# ------------------------------------------------------------------------------------------------
def _fuzz_cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        planes = int(rng.integers(1, 8))
        h = int(rng.integers(2, 90))
        w = int(rng.integers(1, 40)) * 4 if rng.random() < 0.7 else int(rng.integers(2, 150))   # 16-byte rows: TMA tiles
        fy, fx = rng.uniform(0.4, 2.5, 2)
        out.append((planes, h, w, max(1, int(round(h * fy))), max(1, int(round(w * fx)))))
    return out


@pytest.mark.parametrize("case", _fuzz_cases(24, 2024))
def test_resample_shape_fuzz(Fn, case):
    planes, h, w, h1, w1 = case
    rng = np.random.default_rng(planes * 1000003 + h * 1009 + w)
    img = (rng.random((planes, h, w)) * 255).astype(np.float32)
    x = cu(img)
    full = lambda a: np.asarray(a).reshape(planes, h1, w1)
    # R1 rect -> hex
    r_bil = full(O.rect_to_hex_resample(img, (h1, w1), "bilinear"))
    same(Fn.rect_to_hex(x, (h1, w1), "bilinear"), r_bil)                                        # exact: bit-identical float64
    close(Fn.rect_to_hex(x, (h1, w1), "bilinear", out_dtype=torch.float32, math="fast"), r_bil, 255.0)
    if h > 1 and w > 1:
        same(Fn.rect_to_hex(x, (h1, w1), "nearest"), full(O.rect_to_hex_resample(img, (h1, w1), "nearest")))
    # R2 / R4 hex -> rect, hexresize
    h_lin = full(O.hex_to_rect_resample(img, (h1, w1), "linear", twin="np"))
    same(Fn.hex_to_rect(x, (h1, w1), "linear", twin="np"), h_lin)
    close(Fn.hex_to_rect(x, (h1, w1), "linear", out_dtype=torch.float32, math="fast", twin="np"), h_lin, 255.0)
    same(Fn.hex_to_rect(x, (h1, w1), "nearest", twin="np"), full(O.hex_to_rect_resample(img, (h1, w1), "nearest", twin="np")))
    z_lin = full(O.hexresize(img, (h1, w1), "linear"))
    same(Fn.hex_resize(x, (h1, w1), "linear"), z_lin)
    close(Fn.hex_resize(x, (h1, w1), "linear", out_dtype=torch.float32, math="fast"), z_lin, 255.0)
