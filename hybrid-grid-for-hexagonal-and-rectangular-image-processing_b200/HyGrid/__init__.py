"""HyGrid on B200: the reference's Python API surface for the rect<->hex resampling and
hex-lattice filtering path, backed by hand-written sm_100a CUDA kernels (libhygrid_b200.so).

Sub-modules mirror the reference package: ``Image`` (IMAGE), ``HexImage`` (HEXIMAGE),
``geometry_np``, ``geometry_torch``, ``HexFrames``, ``HexModules``; ``functional`` adds the
batched device-resident entry points.  Nothing here falls back to a CPU implementation.
"""
__version__ = "0.1.0"
