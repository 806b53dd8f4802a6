"""Optional raster file back-ends of the container classes (GDAL / mmcv / cv2), resolved lazily.

File I/O is outside the hot path (OUT OF SCOPE) -- none of these packages is needed to resample or
filter, and none is installed on the GPU boxes.  The reference imports all three at module import time and
``sys.exit()``s when one is missing (Image.py:4-27, HexImage.py:5-40); here a missing back-end is an ImportError raised
by the one call that needs it."""
from __future__ import annotations

import importlib
import os

import numpy as np

RASTER_SUFFIXES = (".tif", ".TIF", ".tiff", ".TIFF", ".jpg", ".png", ".jpeg", ".JPEG")     # Image.py:49


def backend(name: str):
    try:
        return importlib.import_module(name)
    except ImportError as e:  # pragma: no cover - depends on the host
        raise ImportError(f"{name} is required for file I/O / display in HyGrid but is not installed") from e


def require_file(path: str) -> str:
    if not os.path.exists(path):
        raise OSError("path dosen't exist.")            # the reference's message (Image.py:47)
    return path


class RasterFile:
    """What ``IMAGE`` keeps as ``self.data`` for an opened raster: size, georeference and windowed reads."""

    def __init__(self, path: str):
        gdal = backend("osgeo.gdal")
        self.handle = gdal.Open(path)
        self.width, self.height, self.bands = self.handle.RasterXSize, self.handle.RasterYSize, self.handle.RasterCount
        self.geotrans, self.proj = self.handle.GetGeoTransform(), self.handle.GetProjection()

    @property
    def shape(self):
        return (self.bands, self.height, self.width)

    def ReadAsArray(self, x0, y0, w, h):
        return self.handle.ReadAsArray(x0, y0, w, h)


def write_raster(path: str, chw: np.ndarray, which: str, geotrans=None, proj=None) -> None:
    """Write a (C, H, W) uint8 / uint16 array with the chosen back-end (HexImage.py:189-213)."""
    if which == 'gdal':
        gdal = backend("osgeo.gdal")
        kind = gdal.GDT_UInt16 if chw.dtype == np.uint16 else gdal.GDT_Byte
        ds = gdal.GetDriverByName("GTiff").Create(path, chw.shape[2], chw.shape[1], chw.shape[0], kind,
                                                  options=["TILED=YES", "COMPRESS=LZW"])
        if geotrans is not None:
            ds.SetGeoTransform(geotrans)
        if proj != None:  # noqa: E711
            ds.SetProjection(proj)
        for band in range(chw.shape[0]):
            ds.GetRasterBand(band + 1).WriteArray(chw[band])
        ds.FlushCache()
        return ds
    hwc = chw[::-1, ...].transpose(1, 2, 0)              # both write BGR, channel-last
    if which == 'mmcv':
        backend("mmcv").imwrite(hwc, path)
    elif which == 'cv2':
        backend("cv2").imwrite(path, hwc)
    return None
