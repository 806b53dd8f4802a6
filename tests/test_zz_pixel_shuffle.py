"""HexPixelShuffle (retired from the reference into "codes in old versions.txt":68-126).

tests/golden/retired_golden.npz holds outputs of the reference's own class (tests/golden/make_retired_golden.py).
CPU: the oracle restatement and the product's host-built index table both reproduce them exactly (index shuffle:
bit-exact).  GPU: the module through ``hg_plane_gather`` / ``hg_plane_scatter`` returns the fixture, equals the oracle
on ragged shapes and dtypes, and its backward is the exact adjoint."""
import os

import numpy as np
import pytest
import torch

from oracle import hexframes_oracle as HO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "retired_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _apply_table(x, r):
    """numpy gather through the product's host table (the exact offsets the kernel is handed)."""
    from HyGrid import HexFrames as hf
    B, C, H, W = x.shape
    cout = C // (r * r)
    n, row, col = hf.pixel_shuffle_table(r, H, W)
    off = np.where(n >= 0, n * (cout * H * W) + row * W + col, -1).reshape(-1)
    flat = x.reshape(B, -1)
    out = np.zeros((B, cout, off.size), np.float32)
    for c in range(cout):
        idx = np.maximum(off, 0) + c * H * W
        out[:, c] = np.where(off >= 0, flat[:, idx], 0)
    return out.reshape((B, cout) + n.shape)


def test_oracle_and_host_table_reproduce_the_reference(golden):
    G = golden
    assert int(G["count"]) >= 10
    for k in range(int(G["count"])):
        r, x, want = int(G[f"ps_{k}_r"]), G[f"ps_{k}_in"], G[f"ps_{k}_out"]
        got = HO.hex_pixel_shuffle(torch.from_numpy(x), r).numpy()
        assert got.dtype == want.dtype == np.float32 and got.shape == want.shape and np.array_equal(got, want)
        tab = _apply_table(x, r)
        assert tab.shape == want.shape and np.array_equal(tab, want)


def test_host_table_on_ragged_shapes_matches_the_oracle():
    from HyGrid import HexFrames as hf
    rng = np.random.default_rng(11)
    for r in (2, 3, 4, 5):
        for H, W in [(1, 1), (1, 4), (2, 1), (3, 2), (5, 7), (8, 3)]:
            x = rng.standard_normal((1, 2 * r * r, H, W)).astype(np.float32)
            want = HO.hex_pixel_shuffle(torch.from_numpy(x), r).numpy()
            assert want.shape == (1, 2, r * H - r + 1, r * W - (r + 1) // 2)
            assert np.array_equal(_apply_table(x, r), want)
            n, _, _ = hf.pixel_shuffle_table(r, H, W)
            assert (n >= 0).all()                                   # every output cell has a source


def test_module_contract_without_a_gpu():
    from HyGrid import HexFrames as hf
    m = hf.HexPixelShuffle(3)
    assert m.upscale_factor == 3 and "3" in repr(m)
    with pytest.raises(ValueError):
        hf.HexPixelShuffle(1)
    with pytest.raises(Exception):
        m(torch.zeros(1, 10, 4, 4))                                 # 10 is not a multiple of 9 (old versions :81-82)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 9, 4, 4))                              # CPU tensor: no CPU path


def test_gather_abi_validates_before_launching():
    from HyGrid import _native as nv
    L = nv.lib()
    assert L.hg_plane_gather(None, None, None, -1, 1, 4, 4, 4, nv.F32, nv.F32, None) == -3
    assert L.hg_plane_gather(None, None, None, 1, 1, 4, 4, 4, nv.F32, nv.F32, None) == -1 and b"null" in L.hg_last_error()
    assert L.hg_plane_gather(None, None, None, 0, 1, 4, 4, 4, nv.F32, nv.F32, None) == 0          # empty batch: nothing to do
    assert L.hg_plane_scatter(None, None, None, 1, 1, 4, -4, 4, nv.F32, None) == -3


@pytest.mark.gpu
def test_gpu_module_returns_the_reference_fixture(golden):
    from HyGrid import HexFrames as hf
    G = golden
    for k in range(int(G["count"])):
        r, x, want = int(G[f"ps_{k}_r"]), G[f"ps_{k}_in"], G[f"ps_{k}_out"]
        got = hf.HexPixelShuffle(r)(torch.from_numpy(x).cuda())
        assert got.dtype == torch.float32 and tuple(got.shape) == want.shape
        assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.gpu
def test_gpu_ragged_shapes_dtypes_and_adjoint():
    from HyGrid import HexFrames as hf
    torch.manual_seed(3)
    for r, cout, B, H, W in [(2, 3, 2, 5, 7), (3, 2, 1, 8, 3), (4, 1, 3, 1, 4), (5, 2, 2, 6, 5), (2, 64, 4, 33, 31)]:
        x = torch.randn(B, cout * r * r, H, W)
        want = HO.hex_pixel_shuffle(x, r)
        m = hf.HexPixelShuffle(r)
        xg = x.cuda().requires_grad_(True)
        y = m(xg)
        assert torch.equal(y.cpu(), want)
        # backward == adjoint of the gather: g_x is exactly the scatter of g through the table, zero where the crop
        # (or a later write) dropped the element; autograd through the oracle's canvas writes is the checker
        g = torch.randn_like(y)
        y.backward(g)
        assert float((y.detach() * g).sum()) == pytest.approx(float((xg.grad * xg.detach()).sum()), rel=1e-4, abs=1e-3)
        if x.numel() <= 4096:                       # (autograd through thousands of canvas writes is slow on big cases)
            xo = x.clone().requires_grad_(True)
            HO.hex_pixel_shuffle(xo, r).backward(g.cpu())
            assert torch.equal(xg.grad.cpu(), xo.grad)
        for dt in (torch.float64, torch.bfloat16, torch.uint8):
            xi = (x * 20).to(dt)
            assert torch.equal(m(xi.cuda()).cpu(), HO.hex_pixel_shuffle(xi, r))
    # 3-D input is treated as one image, like the reference's unsqueeze loop (old versions :73-74)
    x3 = torch.randn(8, 4, 5)
    assert torch.equal(hf.HexPixelShuffle(2)(x3.cuda()).cpu(), HO.hex_pixel_shuffle(x3, 2))
