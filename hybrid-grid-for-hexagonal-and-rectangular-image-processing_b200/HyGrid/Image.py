"""``HyGrid.Image.IMAGE`` -- the rectangular image container of the reference (Image.py:39-159) as the entry of the
rect -> hex path: ``ConvertToHexagon`` hands the pixel array to the rect->hex kernel (C ABI host ring).

Same constructor signature, attributes (``Image, bands, height, width, shape, geotrans, proj, path, backend``),
methods and error messages as the reference; raster files go through the lazy back-ends of ``_rasterio``."""
from __future__ import annotations

import os

import numpy as np

from . import _rasterio as rio
from .geometry_np import rect_to_hex_resample

__all__ = ["IMAGE"]

_IDENTITY_GEOTRANS = (0, 1, 0, 0, 0, 1)


def _one_source(pathname, data):
    """Exactly one of ``pathname`` / ``data`` (Image.py:41-44)."""
    if pathname is None and data is None:
        raise ValueError("pathname and data can not be None at the same time")
    if pathname is not None and data is not None:
        raise ValueError("pathname and data can not be Given at the same time")
    return pathname is not None


def _as_bands(array):
    """A 2-D array is one band (Image.py:61-63): a broadcast view, no copy."""
    return np.broadcast_to(array, (1,) + array.shape) if array.ndim == 2 else array


class IMAGE:
    def __init__(self, pathname=None, data=None, geotrans=None, proj=None, backend='gdal'):
        if _one_source(pathname, data):
            self._from_file(pathname)
        else:
            self._from_array(data, geotrans, proj)
        self.shape = (self.bands, self.height, self.width)
        self.backend = backend

    # -- construction ---------------------------------------------------------------------------
    def _from_array(self, data, geotrans, proj):
        self.Image = _as_bands(data)
        self.bands, self.height, self.width = self.Image.shape
        self.geotrans = _IDENTITY_GEOTRANS if geotrans == None else geotrans  # noqa: E711  (the reference's test)
        self.proj = proj
        self.path = 'tmp.tif'

    def _from_file(self, pathname):
        self.path = rio.require_file(pathname)
        if os.path.splitext(pathname)[1] in rio.RASTER_SUFFIXES:
            self.filetype = 1
            self.data = rio.RasterFile(self.path)
            self.bands, self.height, self.width = self.data.shape
            self.geotrans, self.proj = self.data.geotrans, self.data.proj
        self.Image = _as_bands(self.LoadImageArray())

    # -- reference surface ----------------------------------------------------------------------
    def size(self, index):
        return self.data.shape[index]          # as in the reference this reads the opened file, not the array (Image.py:74-75)

    def Tiles(self):
        """Streaming tiles are announced but not implemented in the reference either (Image.py:81-88)."""

    def LoadImageArray(self, w_range_start=0, h_range_start=0, w_range=None, h_range=None):
        w_range = self.width if w_range is None else w_range
        h_range = self.height if h_range is None else h_range
        window = self.data.ReadAsArray(w_range_start, h_range_start, w_range, h_range)
        self.width, self.height = w_range - w_range_start, h_range - h_range_start
        return np.expand_dims(window, axis=0) if self.bands == 1 else window

    def ConvertToHexagon(self, interpolation='nearest'):
        """Half-resolution hex lattice (Image.py:111-116) -- one pass of the rect->hex kernel."""
        return rect_to_hex_resample(self.Image, [self.height // 2, self.width // 2], interpolation=interpolation)

    def SaveImage(self, pathname):
        if self.backend == 'gdal':
            # the reference never gets past its own driver lookup: drivername is None for every suffix (Image.py:129-134)
            raise Exception("class IMAGE in HyGrid/Image.py: format of output is incorrect, the gdal drivername = None")
        rio.write_raster(pathname, np.asarray(self.Image), self.backend)

    def imshow(self):
        plt = rio.backend("matplotlib.pyplot")
        picture = self.Image.astype(np.uint8)
        plt.imshow(picture.squeeze(), cmap='gray') if self.bands == 1 else plt.imshow(picture.transpose(1, 2, 0)[..., :3])
        plt.show()
