"""The legacy numba twin (reference HyGrid/geometry.py) -- module ``HyGrid.geometry`` here.

tests/golden/numba_twin_golden.npz holds outputs of the UNMODIFIED reference kernel run under numba's CUDA simulator
(tests/golden/make_numba_golden.py).  CPU part: the oracle reproduces them (so the numba twin is pinned to the same
restatement the kernels are checked against) and the host-side lattice builder agrees with the oracle's.
GPU part: ``HyGrid.geometry`` through the C ABI returns the fixture.

Tolerances.  'linear': 1e-12 relative to the value range -- the numba kernel associates ``0.5*i_ + (y_ + (w-.5)/2)``
(geometry.py:28) where the numpy / torch twins compute ``0.5*i_ + y_ + (w-.5)/2`` (geometry_np.py:279), one ulp apart
on some coordinates; every linspace case of the fixture is in fact bit-identical, one warp sample differs by 1.4e-14.
'nearest': exact on the tie-free cases; in the others the only differing samples are exact ``d1 == d3`` distance ties
(documented deviation, see HyGrid/geometry.py), where the fixture value must still be one of the image's own samples.
"""
import os

import numpy as np
import pytest
import torch

from oracle import hygrid_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "numba_twin_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _names(G, kind):
    return sorted({k.split("/")[1] for k in G.files if k.startswith(kind + "/")})


TIE_FREE = {"resample": {"up2", "odd", "same"}, "warp": {"identity", "rotate"}}
MAX_TIES = 0.15          # fraction of samples that may sit on an exact tie in the remaining cases


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b), float(np.abs(a - b).max())


def _check(kind, name, interp, got, want, img):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape and got.dtype == want.dtype == np.float64, (got.shape, want.shape, got.dtype)
    if interp == "linear":
        assert np.abs(got - want).max() <= 1e-12 * 255
    elif name in TIE_FREE[kind]:
        assert np.array_equal(got, want)
    else:
        bad = got != want
        assert bad.mean() <= MAX_TIES and np.isin(want[bad], np.append(np.asarray(img).ravel(), 0.0)).all()


def test_oracle_reproduces_the_numba_resampler(golden):
    G = golden
    assert len(_names(G, "resample")) == 5
    for n in _names(G, "resample"):
        img, size = G[f"resample/{n}/in"], tuple(int(v) for v in G[f"resample/{n}/size"])
        for interp in ("linear", "nearest"):
            want = G[f"resample/{n}/{interp}"]                       # always float64 (C, h1, w1), never squeezed
            got = np.asarray(O.hex_to_rect_resample(img, size, interp, twin="np"), dtype=np.float64).reshape(want.shape)
            _check("resample", n, interp, got, want, img)


def test_oracle_reproduces_the_numba_warp(golden):
    G = golden
    assert len(_names(G, "warp")) == 3
    for n in _names(G, "warp"):
        img, H = G[f"warp/{n}/in"], G[f"warp/{n}/H"]
        for interp in ("linear", "nearest"):
            want = G[f"warp/{n}/{interp}"]
            got = np.asarray(O.hex_warp(img, H, interp, twin="numba"), dtype=np.float64).squeeze()
            _check("warp", n, interp, got, want, img)


def test_numba_tie_rule_is_the_documented_deviation():
    """d1 == d3 < d2 (sample half-way between the first and the third corner): the numba kernel's cascaded ifs keep the
    LAST candidate (geometry.py:131-136), torch.min -- and this build -- the first.  Restated here in plain Python."""
    d1, d2, d3 = 0.3125, 0.8125, 0.3125
    picked = None
    if d1 <= d2 and d1 <= d3:
        picked = 1
    if d2 < d1 and d2 <= d3:
        picked = 2
    if d3 <= d1 and d3 < d2:
        picked = 3
    assert picked == 3 and int(np.argmin([d1, d2, d3])) + 1 == 1


def test_host_lattice_builder_matches_the_oracle():
    from HyGrid import functional as Fn
    t = np.deg2rad(33.0)
    mats = [np.eye(3), np.diag([1.7, 0.6, 1.0]), np.array([[np.cos(t), -np.sin(t), 1.2], [np.sin(t), np.cos(t), -0.4], [0, 0, 1.0]])]
    for h, w in [(9, 8), (12, 17), (5, 30)]:
        for H in mats:
            for twin in ("np", "numba", "torch"):
                rows, cols, Hi = Fn.warp_lattice(h, w, H, twin)
                X, Y, Hi_o = O.warp_coords(h, w, H, twin)
                assert X.shape == (rows.shape[0], cols.shape[0])
                assert np.array_equal(X[:, 0], rows) and np.array_equal(Y[0], cols) and np.array_equal(Hi, Hi_o)
                cx, cy = Fn._warp_planes(rows, cols, Hi, twin)
                hom = np.stack([X, Y, np.ones_like(X)], 0)
                if twin == "torch":
                    inv = torch.einsum("ij, jkl -> ikl", torch.tensor(Hi_o), torch.tensor(hom)).to(torch.float).numpy()
                else:
                    inv = np.einsum("ij, jkl -> ikl", Hi_o, hom)
                assert np.array_equal(cx.numpy(), inv[0]) and np.array_equal(cy.numpy(), inv[1])


def test_module_surface():
    from HyGrid import geometry as g
    for name in ("image_geometric_transformation_gpu", "image_geometric_transformation_cpu", "image_geometric_transformation",
                 "hex_to_square_resample", "hexresize"):
        assert callable(getattr(g, name))
    img = np.zeros((2, 6, 6))
    with pytest.raises(KeyError):
        g.hex_to_square_resample(img, None, "cubic")
    with pytest.raises(NotImplementedError):
        g.hex_to_square_resample(img, None, "bilinear")
    with pytest.raises(NotImplementedError):
        g.image_geometric_transformation(img, np.eye(3), "linear", device="cpu")
    with pytest.raises(Exception):
        g.hex_to_square_resample(np.zeros((1, 2, 3, 4)), None, "linear")


@pytest.mark.gpu
def test_gpu_module_returns_the_numba_fixture(golden):
    from HyGrid import geometry as g
    G = golden
    for n in _names(G, "resample"):
        img, size = G[f"resample/{n}/in"], tuple(int(v) for v in G[f"resample/{n}/size"])
        for interp in ("linear", "nearest"):
            _check("resample", n, interp, g.hex_to_square_resample(img, size, interp), G[f"resample/{n}/{interp}"], img)
    for n in _names(G, "warp"):
        img, H = G[f"warp/{n}/in"], G[f"warp/{n}/H"]
        for interp in ("linear", "nearest"):
            out = g.image_geometric_transformation_gpu(img, H, interp)
            _check("warp", n, interp, out, G[f"warp/{n}/{interp}"], img)
            _same(g.image_geometric_transformation(img, H, interp, device="cuda0"), out)
            _same(out, np.asarray(O.hex_warp(img, H, interp, twin="numba"), dtype=np.float64).squeeze())   # kernel == oracle, exactly


@pytest.mark.gpu
def test_gpu_hexresize_follows_the_numpy_twin():
    from HyGrid import geometry as g
    from HyGrid import geometry_np as gnp
    img = np.random.default_rng(5).random((3, 11, 13)) * 255
    out = g.hexresize(img, (17, 9), "linear")
    assert out.shape == (3, 17, 9) and out.dtype == np.float64
    _same(out, gnp.hexresize(img, (17, 9), "linear"))
    _same(out, np.asarray(O.hexresize(img, (17, 9), "linear")).reshape(out.shape))
