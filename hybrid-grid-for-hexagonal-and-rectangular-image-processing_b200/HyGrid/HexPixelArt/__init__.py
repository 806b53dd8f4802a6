"""``HyGrid.HexPixelArt`` on B200: the hex-mosaic *preview rasteriser* (SURVEY.md section 8f rank 4).

The reference package is an interactive OpenGL viewer (GLFW window, texture upload, a GLSL fragment shader that maps
every screen pixel to the hex cell it falls in: HexPixelArt/hexagon_mosaic_shader.py:25-81).  The window, the event loop
and the texture object have no place on a headless GPU node; what is kept is the shader's pixel -> hex-cell rule,
evaluated into a rect raster by one gather launch (``hexagon_mosaic``).
"""
from .hexagon_mosaic import hexagon_mosaic, mosaic_table

__all__ = ["hexagon_mosaic", "mosaic_table"]
