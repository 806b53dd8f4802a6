"""CPU-side behaviour of the container classes that needs no kernel: constructors, attributes, error types,
type1 / type2 decode (numpy slicing, HexImage.py:65-84, 106-111) and the '.heximg' pickle round trip
(HexImage.py:171-218).  ref: Image.py:39-72, HexImage.py:43-125."""
import numpy as np
import pytest

from HyGrid.HexImage import HEXIMAGE
from HyGrid.Image import IMAGE


def test_image_constructor_contract():
    with pytest.raises(ValueError):
        IMAGE()
    with pytest.raises(ValueError):
        IMAGE(pathname="a.tif", data=np.zeros((3, 4, 4)))
    a = np.arange(24, dtype=np.uint8).reshape(2, 3, 4)
    im = IMAGE(data=a)
    assert (im.bands, im.height, im.width) == (2, 3, 4) and im.shape == (2, 3, 4)
    assert im.geotrans == (0, 1, 0, 0, 0, 1) and im.proj is None
    assert np.array_equal(im.Image, a)
    with pytest.raises(AttributeError):                   # reference quirk kept: size() reads the GDAL dataset handle (Image.py:74-75)
        im.size(1)
    g = IMAGE(data=np.zeros((5, 6)))                      # 2-D input is one band
    assert (g.bands, g.height, g.width) == (1, 5, 6)


def test_heximage_decode_and_attributes():
    with pytest.raises(ValueError):
        HEXIMAGE()
    rng = np.random.default_rng(0)
    t1 = rng.random((3, 6, 11))                           # type1 raster: H x (2W+1)
    h1 = HEXIMAGE(data=t1, heximagetype=1, even_odd_offset=True)
    assert np.array_equal(h1.HexagonImage, t1[:, :, 1:-1:2]) and h1.shape == (3, 6, 5) and h1.even_odd_offset == 1
    t2 = rng.random((2, 8, 9))                            # type2 raster: 2H x (2W+1)
    h2 = HEXIMAGE(data=t2, heximagetype=2)
    assert np.array_equal(h2.HexagonImage, t2[:, ::2, 1:-1:2]) and h2.shape == (2, 4, 4) and h2.even_odd_offset == 0
    h0 = HEXIMAGE(data=rng.random((7, 5)))
    assert h0.shape == (1, 7, 5) and h0.heximagetype is None and h0.geotrans == (0, 1, 0, 0, 0, 1)


def test_heximg_pickle_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    data = (rng.random((3, 9, 7)) * 255).astype(np.float32)
    hx = HEXIMAGE(data=data, geotrans=(10.0, 0.5, 0, 20.0, 0, -0.5), proj="EPSG:4326", even_odd_offset=True)
    path = str(tmp_path / "tile.heximg")
    hx.SaveHexImage(path)
    back = HEXIMAGE(pathname=path)
    assert np.array_equal(back.HexagonImage, data) and back.HexagonImage.dtype == data.dtype
    assert back.shape == (3, 9, 7) and back.even_odd_offset == 1
    assert back.geotrans == (10.0, 0.5, 0, 20.0, 0, -0.5) and back.proj == "EPSG:4326"
    with pytest.raises(OSError):
        HEXIMAGE(pathname=str(tmp_path / "missing.heximg"))
