"""Pin the CPU oracle (oracle/) against outputs of the reference itself
(tests/golden/*.npz, written by tests/golden/make_golden.py from /root/reference).

Integer / nearest / layout results and all float64 interpolation results must be
bit-identical; hex conv is compared at 1e-5 (the reference sums a zero-stuffed
dense window through oneDNN, a different summation order)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import hygrid_oracle as O
from oracle import hexframes_oracle as HO


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b, equal_nan=True), float(np.nanmax(np.abs(a.astype(np.float64) - b)))


def index_image(h, w):
    ii, jj = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    return np.stack([ii + 1, jj + 1], 0).astype(np.float64)


def _dsize(a):
    return None if a[0] < 0 else (int(a[0]), int(a[1]))


def test_r1_rect_to_hex_bit_exact(resample_golden):
    G = resample_golden
    for n in range(int(G["r1_count"])):
        out = O.rect_to_hex_resample(G[f"r1_{n}_img"], _dsize(G[f"r1_{n}_dsize"]), str(G[f"r1_{n}_interp"]))
        same(out, G[f"r1_{n}_out"])


def test_r1_index_tables(resample_golden):
    G = resample_golden
    for n in range(int(G["r1idx_count"])):
        h, w, h1, w1 = (int(v) for v in G[f"r1idx_{n}_hw"])
        out = O.rect_to_hex_resample(index_image(h, w), (h1, w1), "nearest").astype(np.int32)
        assert sha(out) == str(G[f"r1idx_{n}_sha"])
        xs, ys = O.rect2hex_coords(h, w, h1, w1)
        i_n, _, j_n, _ = O.rect2hex_index(h, w, xs, ys)
        # closed form: nearest == truncation sampling, zero outside
        exp_i = np.where((i_n >= 0) & (i_n < h), i_n + 1, 0)
        exp_j = np.where((j_n >= 0) & (j_n < w), j_n + 1, 0)
        assert np.array_equal(exp_i, G[f"r1idx_{n}_i"])
        assert np.array_equal(exp_j, G[f"r1idx_{n}_j"])


def test_config1_round_trip(resample_golden):
    G = resample_golden
    img = np.random.default_rng(0).integers(0, 256, (3, 512, 512), dtype=np.uint8)
    hexd = O.rect_to_hex_resample(img, [256, 256], "nearest")
    assert hexd.shape == tuple(G["c1_hex_shape"]) and str(hexd.dtype) == str(G["c1_hex_dtype"])
    assert sha(hexd) == str(G["c1_hex_sha"])
    back = O.hex_to_rect_resample(hexd, (512, 512), "linear")
    assert sha(back) == str(G["c1_back_sha"])
    same(back[:, ::37, ::41], G["c1_back_probe"])


def test_r2_r4_hex_source_bit_exact(resample_golden):
    G = resample_golden
    for n in range(int(G["r2_count"])):
        img, ds = G[f"r2_{n}_img"], _dsize(G[f"r2_{n}_dsize"])
        same(O.hex_to_rect_resample(img, ds, "linear", twin="np"), G[f"r2_{n}_np_linear"])
        same(O.hex_to_rect_resample(img, ds, "linear", twin="torch"), G[f"r2_{n}_torch_linear"])
        same(O.hex_to_rect_resample(img, ds, "nearest", twin="torch"), G[f"r2_{n}_torch_nearest"])
        same(O.hexresize(img, ds if ds else img.shape[1:], "linear"), G[f"r2_{n}_resize_linear"])


def test_r2_index_tables_torch_twin(resample_golden):
    G = resample_golden
    for n in range(int(G["r2idx_count"])):
        h, w, h1, w1 = (int(v) for v in G[f"r2idx_{n}_hw"])
        out = O.hex_to_rect_resample(index_image(h, w), (h1, w1), "nearest", twin="torch").astype(np.int32)
        assert sha(out) == str(G[f"r2idx_{n}_sha"])


def test_r3_warp(resample_golden):
    G = resample_golden
    for n in range(int(G["r3_count"])):
        img, H = G[f"r3_{n}_img"], G[f"r3_{n}_H"]
        same(O.hex_warp(img, H, "linear", twin="np"), G[f"r3_{n}_np_linear"])
        same(O.hex_warp(img, H, "linear", twin="torch"), G[f"r3_{n}_torch_linear"])
        same(O.hex_warp(img, H, "nearest", twin="torch"), G[f"r3_{n}_torch_nearest"])


def test_r3_identity_is_identity():
    img = np.random.default_rng(1).random((3, 9, 7))
    assert np.array_equal(O.hex_warp(img, np.eye(3), "linear", twin="np"), img)


def test_r5_doubled_rasters(resample_golden):
    G = resample_golden
    for n in range(int(G["r5_count"])):
        img, off = G[f"r5_{n}_img"], int(G[f"r5_{n}_off"])
        same(O.hex_to_type1(img, off), G[f"r5_{n}_t1"])
        same(O.hex_to_type2(img, off), G[f"r5_{n}_t2"])
        same(O.type1_to_hex(G[f"r5_{n}_t1"]), G[f"r5_{n}_dec1"])
        same(O.type2_to_hex(G[f"r5_{n}_t2"]), G[f"r5_{n}_dec2"])
        x = torch.tensor(img, dtype=torch.float32)[None]
        same(HO.heximage_to_type1(x, off).numpy(), G[f"r5_{n}_tt1"])
        same(HO.heximage_to_type2(x, off).numpy(), G[f"r5_{n}_tt2"])
        same(HO.type1_to_heximage(HO.heximage_to_type1(x, off), off)[0].numpy(), G[f"r5_{n}_tdec"])


def test_axial_offset_round_trip():
    i = np.arange(0, 50)[:, None]
    j = np.arange(-5, 40)[None, :]
    assert np.array_equal(O.axial_to_offset(i, O.offset_to_axial(i, j)), np.broadcast_to(j, (50, 45)))
    assert np.array_equal(O.offset_to_axial(i, j), j + (i + 1) // 2)


def test_hexconv_forward_backward(hexframes_golden):
    C = hexframes_golden
    assert int(C["conv_count"]) > 40
    for n in range(int(C["conv_count"])):
        r, s, d, pad, off, g, hb = (int(v) for v in C[f"conv_{n}_cfg"])
        x = torch.tensor(C[f"conv_{n}_x"], requires_grad=True)
        w = torch.tensor(C[f"conv_{n}_w"], requires_grad=True)
        b = torch.tensor(C[f"conv_{n}_b"], requires_grad=True) if hb else None
        y = HO.hexconv2d(x, w, b, off, r, s, pad, d, g)
        ref = torch.tensor(C[f"conv_{n}_y"])
        assert y.shape == ref.shape, (n, y.shape, ref.shape)
        torch.testing.assert_close(y, ref, rtol=1e-5, atol=1e-5)
        (y * torch.tensor(C[f"conv_{n}_gy"])).sum().backward()
        torch.testing.assert_close(x.grad, torch.tensor(C[f"conv_{n}_dx"]), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(w.grad, torch.tensor(C[f"conv_{n}_dw"]), rtol=1e-4, atol=1e-4)
        if hb:
            torch.testing.assert_close(b.grad, torch.tensor(C[f"conv_{n}_db"]), rtol=1e-4, atol=1e-4)


def test_hexconv_adaptive_padding(hexframes_golden):
    C = hexframes_golden
    assert int(C["aconv_count"]) >= 3
    for n in range(int(C["aconv_count"])):
        r, s, d = (int(v) for v in C[f"aconv_{n}_cfg"])
        x = torch.tensor(C[f"aconv_{n}_x"])
        pads = HO.adaptive_padding(x.shape[-2], x.shape[-1], r, s, d)
        xp = torch.nn.functional.pad(x, pads)
        y = HO.hexconv2d(xp, torch.tensor(C[f"aconv_{n}_w"]), torch.tensor(C[f"aconv_{n}_b"]), 0, r, s, 0, d, 1)
        torch.testing.assert_close(y, torch.tensor(C[f"aconv_{n}_y"]), rtol=1e-5, atol=1e-5)


def test_hexpool_forward_backward(hexframes_golden):
    C = hexframes_golden
    assert int(C["pool_count"]) >= 18
    for n in range(int(C["pool_count"])):
        kh, kw, sh, sw, pad, ceil, cip = (int(v) for v in C[f"pool_{n}_cfg"])
        x = torch.tensor(C[f"pool_{n}_x"], requires_grad=True)
        y = HO.hexpool2d(x, str(C[f"pool_{n}_method"]), (kh, kw), (sh, sw), pad, ceil_mode=bool(ceil),
                         count_include_pad=bool(cip))
        method = str(C[f"pool_{n}_method"])
        if method == "average":      # summation order of a 6/9-element window differs: 1e-6
            np.testing.assert_allclose(y.detach().numpy(), C[f"pool_{n}_y"], rtol=1e-6, atol=1e-7, equal_nan=True)
        else:
            same(y.detach().numpy(), C[f"pool_{n}_y"])
        (torch.nan_to_num(y) * torch.tensor(C[f"pool_{n}_gy"])).sum().backward()
        np.testing.assert_allclose(x.grad.numpy(), C[f"pool_{n}_dx"], rtol=1e-6, atol=1e-7)


def test_adaptive_and_global_pool(hexframes_golden):
    C = hexframes_golden
    for n in range(int(C["apool_count"])):
        y = HO.hexadaptivepool2d(torch.tensor(C[f"apool_{n}_x"]), int(C[f"apool_{n}_out"]), str(C[f"apool_{n}_method"]))
        np.testing.assert_allclose(y.numpy(), C[f"apool_{n}_y"], rtol=1e-6, atol=1e-7)
        if str(C[f"apool_{n}_method"]) != "average":
            same(y.numpy(), C[f"apool_{n}_y"])
    for m in ("max", "min", "average"):
        np.testing.assert_allclose(HO.hexglobalpool2d(torch.tensor(C[f"gpool_{m}_x"]), m).numpy(), C[f"gpool_{m}_y"], rtol=1e-6, atol=1e-7)
