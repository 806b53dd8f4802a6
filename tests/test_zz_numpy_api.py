"""The numpy-facing drop-in surface on the GPU: ``geometry_np``, ``geometry_torch``, ``IMAGE`` and ``HEXIMAGE`` called
exactly the way a user of the reference calls them (numpy in, numpy out), against the golden outputs of the reference's
own functions (tests/golden/resample_golden.npz) -- values with ``==``, and also the result's shape after ``.squeeze()``
and its dtype.  (tests/test_gpu_resample.py checks the same kernels through the device-tensor API.)"""
import hashlib

import numpy as np
import pytest

from oracle import hygrid_oracle as O

pytestmark = pytest.mark.gpu


def _dsize(a):
    return None if a[0] < 0 else (int(a[0]), int(a[1]))


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b, equal_nan=True), float(np.nanmax(np.abs(a.astype(np.float64) - b)))


def test_geometry_np_against_the_reference_outputs(resample_golden):
    from HyGrid import geometry_np as gnp
    G = resample_golden
    for n in range(int(G["r1_count"])):
        img, ds, interp = G[f"r1_{n}_img"], _dsize(G[f"r1_{n}_dsize"]), str(G[f"r1_{n}_interp"])
        same(gnp.rect_to_hex_resample(img.copy(), ds, interp), G[f"r1_{n}_out"])
    for n in range(int(G["r2_count"])):
        img, ds = G[f"r2_{n}_img"], _dsize(G[f"r2_{n}_dsize"])
        same(gnp.hex_to_rect_resample(img.copy(), ds, "linear"), G[f"r2_{n}_np_linear"])
        same(gnp.hexresize(img.copy(), ds if ds else img.shape[-2:], "linear"), G[f"r2_{n}_resize_linear"])
        # deviation: 'nearest' works here (the reference raises from a np.min unpacking bug) and follows the rule of
        # the working torch twin on float64 coordinates == the oracle
        near = gnp.hex_to_rect_resample(img.copy(), ds, "nearest")
        want = np.asarray(O.hex_to_rect_resample(img, ds, "nearest", twin="np")).squeeze()
        same(near, want.astype(near.dtype))
    for n in range(int(G["r3_count"])):
        same(gnp.image_geometric_transformation(G[f"r3_{n}_img"].copy(), G[f"r3_{n}_H"], "linear"), G[f"r3_{n}_np_linear"])


def test_geometry_torch_against_the_reference_outputs(resample_golden):
    from HyGrid import geometry_torch as gt
    G = resample_golden
    for n in range(int(G["r2_count"])):
        img, ds = G[f"r2_{n}_img"], _dsize(G[f"r2_{n}_dsize"])
        same(gt.hex_to_square_resample(img.copy(), ds, "linear"), G[f"r2_{n}_torch_linear"])
        same(gt.hex_to_square_resample(img.copy(), ds, "nearest"), G[f"r2_{n}_torch_nearest"])
    for n in range(int(G["r3_count"])):
        img, H = G[f"r3_{n}_img"], G[f"r3_{n}_H"]
        same(gt.image_geometric_transformation_gpu(img.copy(), H, "nearest"), G[f"r3_{n}_torch_nearest"])
        same(gt.image_geometric_transformation_gpu(img.copy(), H, "linear"), G[f"r3_{n}_torch_linear"])
        same(gt.image_geometric_transformation(img.copy(), H, "linear", device="cuda0"), G[f"r3_{n}_torch_linear"])


def test_error_behaviour_of_the_numpy_api():
    from HyGrid import geometry_np as gnp
    from HyGrid import geometry_torch as gt
    img = np.zeros((3, 8, 8), np.float32)
    for fn in (gnp.hex_to_rect_resample, gnp.hexresize, gt.hex_to_square_resample):
        with pytest.raises(NotImplementedError):
            fn(img, (8, 8), "bilinear")                       # the reference returns uninitialised memory here
        with pytest.raises(KeyError):
            fn(img, (8, 8), "cubic")
    with pytest.raises(KeyError):
        gnp.rect_to_hex_resample(img, None, "linear")         # rect->hex knows 'nearest' and 'bilinear' only
    for fn in (gnp.rect_to_hex_resample, gnp.hex_to_rect_resample):
        with pytest.raises(Exception):
            fn(np.zeros((2, 2, 2, 2)), None, "nearest")
    assert gnp.rect_to_hex_resample(np.zeros((8, 8), np.uint8), (4, 4), "nearest").shape == (4, 4)   # 2-D = one band, squeezed


def test_image_and_heximage_classes(resample_golden):
    from HyGrid.HexImage import HEXIMAGE
    from HyGrid.Image import IMAGE
    G = resample_golden
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()   # noqa: E731
    img = np.random.default_rng(0).integers(0, 256, (3, 512, 512), dtype=np.uint8)   # BASELINE config 1
    hexd = IMAGE(data=img).ConvertToHexagon()
    assert isinstance(hexd, np.ndarray) and hexd.dtype == np.uint8 and hexd.shape == (3, 256, 256)
    assert sha(hexd) == str(G["c1_hex_sha"])
    for n in range(int(G["r5_count"])):
        src, off = G[f"r5_{n}_img"], int(G[f"r5_{n}_off"])
        hx = HEXIMAGE(data=src.copy(), even_odd_offset=bool(off))
        t1, g1 = hx.GenerateType1Image()
        t2, g2 = hx.GenerateType2Image()
        same(t1, G[f"r5_{n}_t1"])
        same(t2, G[f"r5_{n}_t2"])
        assert g1[5] == 2 * hx.geotrans[5] and g2 == tuple(hx.geotrans)
        same(HEXIMAGE(data=t1, heximagetype=1).HexagonImage, G[f"r5_{n}_dec1"])
        same(HEXIMAGE(data=t2, heximagetype=2).HexagonImage, G[f"r5_{n}_dec2"])
