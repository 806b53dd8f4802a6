"""``HyGrid.HexImage.HEXIMAGE`` -- the hex image container of the reference (HexImage.py:43-276) on the formats either
side of the path: type1 / type2 doubled rasters and ``.heximg`` pickles.

* decode (constructor, ``heximagetype`` 1 / 2): the reference's own strided views (HexImage.py:106-111);
* encode (``GenerateType1Image`` / ``GenerateType2Image``; per-row ``np.insert`` / ``np.append`` loops in the reference,
  :139-170): the pack kernel, fed from and drained into HOST memory by the C ABI's pinned ring
  (``hg_host_hex_to_type``: chunks of bands flow H2D -> kernel -> D2H on three streams) -- no torch tensors involved;
* ``SaveHexImage``: ``.heximg`` pickles are written here, rasters through the lazy back-ends of ``_rasterio``;
* ``Hex_imshow`` needs a display; ``HexMosaic()`` returns the raster its fragment shader would draw.

Same constructor signature, attributes and error messages as the reference."""
from __future__ import annotations

import ctypes as C
import os
import pickle
import warnings

import numpy as np

from . import _native as nv
from . import _rasterio as rio
from .Image import IMAGE, _IDENTITY_GEOTRANS, _as_bands, _one_source
from ._hostapi import device_index

__all__ = ["HEXIMAGE"]

_BAD_TYPE = ("不支持的文件类型\n要么输入的是普通图像文件：None\n要么输入的是六边形图像通用格式：1\n"
             "要么输入后缀为‘.heximg'的六边形图像专用文件格式：2")           # HexImage.py:85-88
# hex matrix inside a doubled raster handed over as an array (HexImage.py:106-111)
_DECODE = {None: lambda a: a, 1: lambda a: a[:, :, 1:-1:2], 2: lambda a: a[:, ::2, 1:-1:2]}
_HEXIMG_KEYS = (("height", "height"), ("width", "width"), ("bands", "bands"), ("geotrans", "geotransform"),
                ("proj", "projection"), ("even_odd_offset", "offset"), ("HexagonImage", "HexMatrix"))


class HEXIMAGE(IMAGE):
    def __init__(self, pathname=None, heximagetype=None, data=None, geotrans=None, proj=None, even_odd_offset=False,
                 backend='gdal'):
        if not _one_source(pathname, data):
            self._from_hex_array(data, heximagetype, geotrans, proj)
        elif os.path.splitext(pathname)[1] == ".heximg":
            even_odd_offset = self._from_heximg(pathname)
        else:
            self._from_raster(pathname, heximagetype, backend)
        self.backend = backend
        self.heximagetype = heximagetype
        self.even_odd_offset = int(even_odd_offset)
        self.shape = (self.bands, self.height, self.width)

    # -- construction ---------------------------------------------------------------------------
    def _from_hex_array(self, data, heximagetype, geotrans, proj):
        self.HexagonImage = _DECODE[heximagetype](_as_bands(data)) if heximagetype in _DECODE else _as_bands(data)
        self.bands, self.height, self.width = self.HexagonImage.shape[:3]
        self.geotrans = _IDENTITY_GEOTRANS if geotrans == None else geotrans  # noqa: E711
        self.proj = proj
        self.path = 'data'                                   # the reference stores the parameter's name (HexImage.py:121)

    def _from_heximg(self, pathname):
        self.path = self.datapath = rio.require_file(pathname)
        with open(pathname, "rb") as f:
            self.Heximagedataset = pickle.load(f)
        self.filetype = 2
        stored = {attr: self.Heximagedataset[key] for attr, key in _HEXIMG_KEYS}
        offset = stored.pop("even_odd_offset")
        self.__dict__.update(stored)
        if self.HexagonImage.ndim < 3:                       # HexImage.py:99-100
            self.HexagonImage = np.broadcast_to(self.HexagonImage, (3, self.height, self.width))
        return offset

    def _from_raster(self, pathname, heximagetype, backend):
        IMAGE.__init__(self, pathname, backend=backend)
        if heximagetype == None:  # noqa: E711                 a plain picture: resample it onto the hex lattice (:61-64)
            hexed = self.ConvertToHexagon()
            self.HexagonImage = hexed[None] if hexed.ndim == 2 else hexed
            self.bands, self.height, self.width = self.HexagonImage.shape[:3]
            return
        if heximagetype not in (1, 2):
            raise Exception(_BAD_TYPE)
        raster = self.LoadImageArray()
        if heximagetype == 2:                                # :72-84: rows were written twice; an even width lost its last zero
            if (self.width & 1) == 0:
                raster = np.append(raster, np.zeros((self.bands, self.height, 1)), axis=2)
                self.width += 1
            self.height //= 2
            raster = raster[:, ::2] if raster.ndim == 3 else raster[::2]
        self.width = (self.width - 1) // 2
        self.HexagonImage = np.zeros([self.bands, self.height, self.width])
        self.HexagonImage[:, :, :] = raster[..., 1::2]

    # -- reference surface ----------------------------------------------------------------------
    def size(self, index):
        return self.HexagonImage.shape[index]

    def build_Heximagedataset(self):
        self.Heximagedataset = {key: getattr(self, attr) for attr, key in _HEXIMG_KEYS}

    def _encode(self, rows_mul):
        """(C, rows_mul*H, 2W+1) float64 doubled raster of the hex matrix: host array in, host array out through the
        library's pinned ring (HexImage.py:139-170)."""
        hexm = np.ascontiguousarray(self.HexagonImage)
        if hexm.dtype not in nv._NP2HG:
            hexm = hexm.astype(np.float64)
        out = np.empty((self.bands, rows_mul * self.height, 2 * self.width + 1), dtype=np.float64)
        nv.call("hg_host_hex_to_type", C.c_void_p(hexm.ctypes.data), C.c_void_p(out.ctypes.data), self.bands, self.height,
                self.width, self.even_odd_offset, nv.hg_dtype(hexm.dtype), nv.F64, rows_mul, device_index())
        return out

    def GenerateType1Image(self):
        """-> (C x H x (2W+1) float64, geotrans with the row pitch doubled) (HexImage.py:139-153)."""
        g = self.geotrans
        return self._encode(1), (g[0], g[1], g[2], g[3], g[4], g[5] * 2,)

    def GenerateType2Image(self):
        """-> (C x 2H x (2W+1) float64, geotrans) (HexImage.py:154-170)."""
        g = self.geotrans
        return self._encode(2), (g[0], g[1], g[2], g[3], g[4], g[5],)

    def SaveHexImage(self, pathname, imagetype=1, filetype=1):
        """HexImage.py:171-218."""
        stem, suffix = os.path.splitext(pathname)
        if suffix == ".heximg":
            filetype = 2
        if suffix in (".tif", ".TIF", ".tiff", ".TIFF", ".png", "bmp"):
            self.filetype = 1
        if suffix in ("JPG", ".jpg", "JPEG", "jpeg"):
            warnings.warn("jpg and jpeg are lossy compression formats, switching to png")
            suffix = ".png"
        pathname = stem + suffix
        if filetype != 1:
            self.build_Heximagedataset()
            with open(pathname, "wb") as f:
                pickle.dump(self.Heximagedataset, f)
            return
        raster, geotrans_out = self.GenerateType1Image() if imagetype == 1 else self.GenerateType2Image()
        raster = raster.astype(np.uint16 if 'int16' in self.HexagonImage.dtype.name else np.uint8)
        handle = rio.write_raster(pathname, raster, self.backend, geotrans_out, self.proj)
        if handle is not None:
            self.Hex_dataset = handle

    def Hex_imshow(self):
        raise NotImplementedError("the interactive OpenGL viewer (HexImage.py:219-276) has no place on a headless GPU node; "
                                  "HexMosaic() returns the raster its fragment shader would draw")

    def HexMosaic(self, out_size=None, hierarchy=0):
        """The picture ``Hex_imshow`` draws, as an array: ``(bands, out_h, out_w)`` raster in which every pixel shows the
        hex cell it falls in (HexPixelArt/hexagon_mosaic_shader.py:25-81 evaluated by one gather launch)."""
        from .HexPixelArt import hexagon_mosaic
        return hexagon_mosaic(self.HexagonImage, out_size, int(self.even_odd_offset), hierarchy)
