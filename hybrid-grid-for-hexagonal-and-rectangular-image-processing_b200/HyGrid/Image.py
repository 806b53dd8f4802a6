"""``HyGrid.Image.IMAGE`` (/root/reference/HyGrid/Image.py:39-159): rectangular image container whose
``ConvertToHexagon`` runs the rect->hex kernel.  File I/O (GDAL / mmcv / cv2) and ``imshow`` are outside
the hot path: the optional back-ends are imported lazily and raise ImportError when used without being
installed (the reference ``sys.exit()``s at import time instead, Image.py:4-27)."""
from __future__ import annotations

import os

import numpy as np

from .geometry_np import rect_to_hex_resample

__all__ = ["IMAGE"]


def _need(name):
    import importlib
    try:
        return importlib.import_module(name)
    except ImportError as e:  # pragma: no cover - depends on the host
        raise ImportError(f"{name} is required for file I/O / display in HyGrid.Image but is not installed") from e


class IMAGE:
    def __init__(self, pathname=None, data=None, geotrans=None, proj=None, backend='gdal'):
        if pathname is None and data is None:
            raise ValueError("pathname and data can not be None at the same time")
        if pathname is not None and data is not None:
            raise ValueError("pathname and data can not be Given at the same time")
        if pathname is not None:
            self.path = pathname
            if not os.path.exists(self.path):
                raise OSError("path dosen't exist.")
            file_name, file_extension = os.path.splitext(pathname)
            if file_extension in (".tif", ".TIF", ".tiff", ".TIFF", ".jpg", ".png", ".jpeg", ".JPEG"):
                gdal = _need("osgeo.gdal")
                self.filetype = 1
                self.data = gdal.Open(self.path)
                self.height = self.data.RasterYSize
                self.width = self.data.RasterXSize
                self.bands = self.data.RasterCount
                self.geotrans = self.data.GetGeoTransform()
                self.proj = self.data.GetProjection()
            self.Image = self.LoadImageArray()
            if self.Image.ndim == 2:
                self.Image = np.broadcast_to(self.Image, (1, self.height, self.width))
        elif data is not None:
            if data.ndim == 2:
                data = np.broadcast_to(data, (1, data.shape[0], data.shape[1]))
            self.Image = data
            self.bands, self.height, self.width = data.shape
            self.geotrans = geotrans
            if self.geotrans == None:  # noqa: E711
                self.geotrans = (0, 1, 0, 0, 0, 1)
            self.proj = proj
            self.path = 'tmp.tif'
        self.shape = (self.bands, self.height, self.width)
        self.backend = backend

    def size(self, index):
        return self.data.shape[index]

    def Tiles(self):
        pass

    def LoadImageArray(self, w_range_start=0, h_range_start=0, w_range=None, h_range=None):
        if w_range is None:
            w_range = self.width
        if h_range is None:
            h_range = self.height
        tmp_image = self.data.ReadAsArray(w_range_start, h_range_start, w_range, h_range)
        self.width = w_range - w_range_start
        self.height = h_range - h_range_start
        if self.bands == 1:
            tmp_image = np.expand_dims(tmp_image, axis=0)
        return tmp_image

    def ConvertToHexagon(self, interpolation='nearest'):
        """Image.py:111-116: half-resolution hex lattice, on the GPU."""
        return rect_to_hex_resample(self.Image, [self.height // 2, self.width // 2], interpolation=interpolation)

    def SaveImage(self, pathname):
        if self.backend == 'gdal':
            # the reference raises here unconditionally (drivername is hard-coded to None, Image.py:129-134)
            raise Exception("class IMAGE in HyGrid/Image.py: format of output is incorrect, the gdal drivername = None")
        elif self.backend == 'mmcv':
            _need("mmcv").imwrite(self.Image[::-1, ...].transpose(1, 2, 0), file_path=pathname)
        elif self.backend == 'cv2':
            _need("cv2").imwrite(pathname, self.Image[::-1, ...].transpose(1, 2, 0))

    def imshow(self):
        plt = _need("matplotlib.pyplot")
        image = self.Image.astype(np.uint8)
        if self.bands == 1:
            plt.imshow(image.squeeze(), cmap='gray')
        else:
            plt.imshow(image.transpose(1, 2, 0)[..., :3])
        plt.show()
