"""ctypes binding of libhygrid_b200.so -- the only way the Python layer reaches the GPU.

There is deliberately no fallback: if the shared library is missing or a call fails the
caller gets an exception, never a silently different (CPU / eager-PyTorch) result.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libhygrid_b200.so")

# element-type enum of include/hygrid_b200.h
U8, I16, I32, I64, F32, F64, BF16, U16 = range(8)
MATH_EXACT, MATH_FAST = 0, 1
POOL_MAX, POOL_MIN, POOL_AVG = 0, 1, 2

_TORCH2HG = {torch.uint8: U8, torch.int16: I16, torch.int32: I32, torch.int64: I64,
             torch.float32: F32, torch.float64: F64, torch.bfloat16: BF16}
_HG2TORCH = {v: k for k, v in _TORCH2HG.items()}
_NP2HG = {np.dtype(np.uint8): U8, np.dtype(np.int16): I16, np.dtype(np.int32): I32, np.dtype(np.int64): I64,
          np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.uint16): U16}


class HyGridNativeError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [("N", C.c_int64), ("Cin", C.c_int64), ("Cout", C.c_int64), ("H", C.c_int64), ("W", C.c_int64),
                ("Ho", C.c_int64), ("Wo", C.c_int64),
                ("radius", C.c_int), ("stride", C.c_int), ("dilation", C.c_int), ("groups", C.c_int),
                ("pad", C.c_int), ("parity", C.c_int), ("pad_value", C.c_float),
                ("x_dtype", C.c_int), ("y_dtype", C.c_int), ("algo", C.c_int), ("relu", C.c_int), ("pad_mode", C.c_int), ("accumulate", C.c_int)]


class TapsSet(C.Structure):
    _fields_ = [("T", C.c_int), ("sx", C.c_int), ("doubled", C.c_int), ("ry", C.c_int * 64), ("ex", C.c_int * 64)]


class DwTapsDesc(C.Structure):
    _fields_ = [("N", C.c_int64), ("C", C.c_int64), ("H", C.c_int64), ("W", C.c_int64), ("Ho", C.c_int64), ("Wo", C.c_int64),
                ("sy", C.c_int), ("odd_limit", C.c_int), ("parity", C.c_int), ("pad", C.c_int), ("pad_value", C.c_float),
                ("set", TapsSet * 2)]


_p, _i, _l, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_SIGS = {
    "hg_axial_to_offset_i32": [_p, _p, _p, _l, _p],
    "hg_offset_to_axial_i32": [_p, _p, _p, _l, _p],
    "hg_rect2hex_index": [_p, _p, _l, _l, _l, _l, _p, _p, _p, _p, _p],
    "hg_hexsrc_index": [_p, _p, _i, _i, _l, _l, _l, _l, _p, _p, _p, _p, _p],
    "hg_rect2hex_nearest": [_p, _p, _p, _p, _l, _l, _l, _l, _l, _i, _p],
    "hg_rect2hex_bilinear": [_p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _l, _i, _i, _i, _p],
    "hg_hex2rect_nearest": [_p, _p, _p, _p, _l, _l, _l, _l, _l, _i, _p],
    "hg_hex2rect_linear": [_p, _p, _p, _p, _p, _p, _l, _l, _l, _l, _l, _i, _i, _i, _p],
    "hg_hexwarp_nearest": [_p, _p, _p, _p, _i, _l, _l, _l, _l, _l, _i, _p],
    "hg_hexwarp_linear": [_p, _p, _p, _p, _i, _l, _l, _l, _l, _l, _i, _i, _p],
    "hg_hexwarp_affine": [_p, _p, C.POINTER(_d), _d, _d, _i, _i, _l, _l, _l, _l, _l, _i, _i, _p],
    "hg_hex_to_type1": [_p, _p, _l, _l, _l, _i, _i, _i, _p],
    "hg_hex_to_type2": [_p, _p, _l, _l, _l, _i, _i, _i, _p],
    "hg_type_to_hex": [_p, _p, _l, _l, _l, _i, _i, _i, _p],
    "hg_pad2d": [_p, _p, _l, _l, _l, _i, _i, _i, _i, _i, _d, _i, _p],
    "hg_pad2d_bwd": [_p, _p, _l, _l, _l, _i, _i, _i, _i, _i, _i, _p],
    "hg_plane_gather": [_p, _p, _p, _l, _l, _l, _l, _l, _i, _i, _p],
    "hg_plane_scatter": [_p, _p, _p, _l, _l, _l, _l, _l, _i, _p],
    "hg_hexpool_fwd": [_p, _p, _p, _i, _l, _l, _l, _l, _l, _i, _i, _i, _i, _i, _i, _d, _i, _i, _d, _i, _i, _p],
    "hg_hexpool_bwd": [_p, _p, _i, _p, _p, _l, _l, _l, _l, _l, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "hg_hexglobalpool_fwd": [_p, _p, _p, _l, _l, _i, _i, _p],
    "hg_hexglobalpool_bwd": [_p, _p, _p, _p, _l, _l, _i, _i, _p],
    "hg_split_bf16": [_p, _p, _p, _l, _p],
    "hg_broadcast_fill": [_p, _p, _l, _i, _p],
    "hg_hexconv_out_shape": [_l, _l, _i, _i, _i, _i, C.POINTER(_l), C.POINTER(_l)],
    "hg_hexconv_umma_eligible": [C.POINTER(ConvDesc), _i],
    "hg_hexconv_fwd": [C.POINTER(ConvDesc), _p, _p, _p, _p, _p],
    "hg_hexconv_fwd_affine": [C.POINTER(ConvDesc), _p, _p, _p, _p, _p, _p],
    "hg_bn_stats": [_p, _p, _l, _l, _l, _p],
    "hg_bn_apply": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _l, _l, C.c_float, _i, _p],
    "hg_bn_bwd_reduce": [_p, _p, _p, _p, _p, _p, _p, _l, _l, _l, _i, _p],
    "hg_bn_bwd_apply": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _l, _l, _i, _i, _p],
    "hg_bn_train_fwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _d, _l, _l, _l, C.c_float, _i, _p],
    "hg_bn_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _l, _l, _l, _i, _i, _p],
    "hg_hexconv_dgrad": [C.POINTER(ConvDesc), _p, _p, _p, _p],
    "hg_hexconv_wgrad": [C.POINTER(ConvDesc), _p, _p, _p, _p, _p],
    "hg_dwtaps_fwd": [C.POINTER(DwTapsDesc), _p, _p, _p, _p, _p],
    "hg_dwtaps_dgrad": [C.POINTER(DwTapsDesc), _p, _p, _p, _p, _p],
    "hg_dwtaps_wgrad": [C.POINTER(DwTapsDesc), _p, _p, _p, _i, _p],
    "hg_host_rect2hex": [_p, _p, _p, _p, _l, _l, _l, _l, _l, _i, _i, _i, _i, _i],
    "hg_host_hex2rect": [_p, _p, _p, _p, _l, _l, _l, _l, _l, _i, _i, _i, _i, _i],
    "hg_host_hex_to_type": [_p, _p, _l, _l, _l, _i, _i, _i, _i, _i],
    "hg_host_type_to_hex": [_p, _p, _l, _l, _l, _i, _i, _i, _i],
}

_lib = None


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HyGridNativeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(or <package>/build.py). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.hg_version.restype = C.c_int
        L.hg_last_error.restype = C.c_char_p
        L.hg_last_launch.restype = C.c_char_p
        L.hg_launch_count.restype = C.c_int64
        L.hg_reset_launch_count.restype = None
        L.hg_host_release.restype = None
        for name, sig in _SIGS.items():
            fn = getattr(L, name, None)
            if fn is None:
                raise HyGridNativeError(f"{LIB_PATH} does not export {name}; rebuild it")
            fn.argtypes = sig
            fn.restype = C.c_int
        _lib = L
    return _lib


_fn_cache = {}


def _fn(name):
    """Bound ctypes function by name (the attribute lookup on a CDLL is the slow part of a call)."""
    f = _fn_cache.get(name)
    if f is None:
        f = _fn_cache[name] = getattr(lib(), name)
    return f


def query(name, *args) -> int:
    """Call an entry point whose return value is an answer, not a status."""
    return int(_fn(name)(*args))


def call(name, *args):
    """Run one entry point; a non-zero status becomes an exception carrying hg_last_error().
    When the stream argument (always last) belongs to another device than the current one, the call is made with
    that device current, so tensors on ``cuda:1`` work from a process whose current device is ``cuda:0``."""
    f = _fn(name)
    st = args[-1] if args else None
    if isinstance(st, StreamArg) and st.device_index != torch.cuda.current_device():
        with torch.cuda.device(st.device_index):
            rc = f(*args)
    else:
        rc = f(*args)
    if rc != 0:
        raise HyGridNativeError(f"{name} failed with code {rc}: {lib().hg_last_error().decode()}")


def last_launch() -> str:
    """Kernel family of this thread's last launch (hg_last_launch)."""
    return lib().hg_last_launch().decode()


def launch_count() -> int:
    return int(lib().hg_launch_count())


def reset_launch_count() -> None:
    lib().hg_reset_launch_count()


def hg_dtype(t) -> int:
    if isinstance(t, torch.dtype):
        try:
            return _TORCH2HG[t]
        except KeyError:
            raise TypeError(f"unsupported tensor dtype {t}") from None
    try:
        return _NP2HG[np.dtype(t)]
    except KeyError:
        raise TypeError(f"unsupported array dtype {t}") from None


def torch_dtype(code: int) -> torch.dtype:
    return _HG2TORCH[code]


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class StreamArg(C.c_void_p):
    """cudaStream_t handle that remembers which device it belongs to (see ``call``)."""
    device_index = -1


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on ``device`` (the raw handle, without building a Stream object: these
    wrappers sit on a host-bound path -- a C5 training step is ~130 launches in 2 ms)."""
    if device is None:
        idx = torch.cuda.current_device()
    else:
        idx = device.index if isinstance(device, torch.device) else torch.device(device).index
        if idx is None:
            idx = torch.cuda.current_device()
    if _raw_stream is not None:
        arg = StreamArg(_raw_stream(idx))
    else:  # pragma: no cover - older torch
        arg = StreamArg(torch.cuda.current_stream(idx).cuda_stream)
    arg.device_index = int(idx)
    return arg


def require_cuda(t: torch.Tensor, what="tensor"):
    if not t.is_cuda:
        raise HyGridNativeError(f"{what} must live on a CUDA device (got {t.device}); there is no CPU path")
    return t
