"""Hex-mosaic preview rasteriser (``HyGrid.HexPixelArt.hexagon_mosaic``).

The reference is a GLSL fragment shader that cannot run without OpenGL (absent here).  Pin: ``tests/golden/make_mosaic_golden.py``
cut the shader SOURCE TEXT out of the reference file, translated it statement by statement (``tests/golden/glsl_mini.py``:
binary32 floats, truncating int division / conversion, implicit int -> float) and executed it for every fragment of six
rasters; ``test_oracle_and_host_table_match_the_reference_shader_text`` holds the oracle (``oracle/hexmosaic_oracle.py``, the
shader restated by hand) and the product's vectorised host table to those texel coordinates.  What stays unpinned is the GL
pipeline around the shader (texture filtering / mip-mapping, texture.py:47-50).  CPU: additionally the product's host table
against the per-fragment oracle cell by cell, plus geometric properties of the rule.  GPU: the gather through the C ABI equals
the oracle raster exactly (pure index shuffle)."""
import numpy as np
import pytest
import torch

from oracle import hexmosaic_oracle as MO

CASES = [(5, 6, 40, 52, 0, 0), (5, 6, 40, 52, 1, 0), (8, 8, 33, 47, 0, 0), (4, 7, 64, 64, 1, 1), (9, 3, 50, 20, 0, 2),
         (1, 1, 9, 9, 0, 0), (12, 16, 24, 32, 1, 0)]     # H, W, out_h, out_w, even_odd_offset, hierarchy


def _oracle_table(H, W, oh, ow, eo, hier):
    th, tw = (H + 3) // 4 * 4, (W + 3) // 4 * 4
    tab = np.full((oh, ow), -1, np.int64)
    for py in range(oh):
        v = (np.float32(py) + np.float32(0.5)) / np.float32(oh)
        for px in range(ow):
            u = (np.float32(px) + np.float32(0.5)) / np.float32(ow)
            r, c = MO.fragment_cell(u, v, tw, th, eo, 2.0 ** (-hier))
            if 0 <= r < H and 0 <= c < W:
                tab[py, px] = r * W + c
    return tab


def test_oracle_and_host_table_match_the_reference_shader_text():
    import os
    from HyGrid.HexPixelArt import mosaic_table
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mosaic_golden.npz"))
    assert int(G["count"]) >= 6 and "even_odd_offset" in str(G["translated_source"])
    for n in range(int(G["count"])):
        th, tw, oh, ow, eo, hier = (int(v) for v in G[f"{n}_cfg"])
        sx, sy = G[f"{n}_sx"], G[f"{n}_sy"]                       # texel coordinate handed to texture2D: a texel centre
        assert np.array_equal(sx - np.floor(sx), np.full_like(sx, 0.5)) and np.array_equal(sy - np.floor(sy), np.full_like(sy, 0.5))
        col, row = np.floor(sx).astype(np.int64), np.floor(sy).astype(np.int64)
        for py in range(oh):
            v = (np.float32(py) + np.float32(0.5)) / np.float32(oh)
            for px in range(ow):
                u = (np.float32(px) + np.float32(0.5)) / np.float32(ow)
                assert MO.fragment_cell(u, v, tw, th, eo, 2.0 ** (-hier)) == (row[py, px], col[py, px]), (n, py, px)
        inside = (row >= 0) & (row < th) & (col >= 0) & (col < tw)
        want = np.where(inside, row * tw + col, -1)
        assert np.array_equal(mosaic_table(th, tw, oh, ow, eo, hier), want), n


def test_host_table_equals_the_per_fragment_oracle():
    from HyGrid.HexPixelArt import mosaic_table
    for H, W, oh, ow, eo, hier in CASES:
        assert np.array_equal(mosaic_table(H, W, oh, ow, eo, hier), _oracle_table(H, W, oh, ow, eo, hier)), (H, W, oh, ow, eo, hier)


def test_rule_properties():
    """Level 0: cell (i, j) is centred at lattice point x = j + 0.5 + 0.5*((i + offset + 1) % 2) ... in texture units --
    checked indirectly: every cell of the lattice is hit, interior cells cover (nearly) equal areas, and a fragment at
    a cell's own texel centre row maps back to that row."""
    from HyGrid.HexPixelArt import mosaic_table
    H, W, s = 8, 8, 24
    for eo in (0, 1):
        tab = mosaic_table(H, W, H * s, W * s, eo, 0)
        hit = np.bincount(tab[tab >= 0], minlength=H * W).reshape(H, W)
        assert (hit > 0).all()
        inner = hit[1:-1, 1:-1].astype(np.float64)
        assert inner.max() / inner.min() < 1.15                    # equal-area cells up to rasterisation noise
        rows = tab // W
        for i in range(1, H - 1):                                  # the texture is (H+1) rows tall in lattice units (:42)
            py = int((i + 1) * (H * s) / (H + 1))                  # lattice y = i + 1 is the centre line of texel row i
            got = rows[py][tab[py] >= 0]
            assert (got == i).all()
    # odd rows are shifted by half a cell: the column boundaries of consecutive rows interleave
    tab = mosaic_table(H, W, H * s, W * s, 0, 0)
    py0, py1 = int(2 * (H * s) / (H + 1)), int(3 * (H * s) / (H + 1))
    b0 = np.flatnonzero(np.diff(tab[py0] % W) != 0)
    b1 = np.flatnonzero(np.diff(tab[py1] % W) != 0)
    assert len(b0) >= W - 2 and len(b1) >= W - 2 and abs(abs(int(b0[2]) - int(b1[2])) - s * W / (W + 0.5) / 2) <= 2


def test_argument_checks_without_a_gpu():
    from HyGrid.HexPixelArt import hexagon_mosaic
    with pytest.raises(Exception):
        hexagon_mosaic(np.zeros(5))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            hexagon_mosaic(np.zeros((3, 4, 4), np.uint8))
        with pytest.raises(RuntimeError):
            hexagon_mosaic(torch.zeros(3, 4, 4))


@pytest.mark.gpu
def test_gpu_mosaic_equals_the_oracle_raster():
    from HyGrid.HexPixelArt import hexagon_mosaic
    rng = np.random.default_rng(2)
    for H, W, oh, ow, eo, hier in CASES:
        img = rng.integers(1, 256, (3, H, W)).astype(np.uint8)
        want = MO.hexagon_mosaic(img, (oh, ow), eo, hier)
        got = hexagon_mosaic(img, (oh, ow), eo, hier)
        assert got.dtype == np.uint8 and np.array_equal(got, want)
        imgf = rng.random((2, 3, H, W)).astype(np.float32)
        gotf = hexagon_mosaic(torch.from_numpy(imgf).cuda(), (oh, ow), eo, hier)
        assert gotf.dtype == torch.float32 and tuple(gotf.shape) == (2, 3, oh, ow)
        for b in range(2):
            assert np.array_equal(gotf[b].cpu().numpy(), MO.hexagon_mosaic(imgf[b], (oh, ow), eo, hier))
    img16 = rng.integers(0, 60000, (1, 6, 6)).astype(np.uint16)        # widened on the host, returned in its own dtype
    out = hexagon_mosaic(img16, (30, 30))
    assert out.dtype == np.uint16 and np.array_equal(out, MO.hexagon_mosaic(img16, (30, 30)))
    assert hexagon_mosaic(np.ones((1, 5, 6), np.float64)).shape == (1, 32, 32)   # default: 4 pixels per padded texel
