"""CPU-side checks of the drop-in boundary: libhygrid_b200.so loads and exports every symbol
include/hygrid_b200.h declares (no compute calls without a GPU), the ctypes table covers the header,
and argument validation that happens before any launch returns the documented error codes."""
import ctypes as C
import os
import re

import pytest

from HyGrid import _native as nv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hygrid_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    assert len(syms) >= 30
    for must in ("hg_rect2hex_bilinear", "hg_hex2rect_linear", "hg_hexwarp_affine", "hg_hexsrc_index",
                 "hg_hex_to_type1", "hg_hexpool_fwd", "hg_hexpool_bwd", "hg_hexconv_fwd", "hg_hexconv_dgrad",
                 "hg_hexconv_wgrad", "hg_host_rect2hex", "hg_host_hex2rect"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(nv.LIB_PATH), "build the library first: python __graft_entry__.py"
    lib = C.CDLL(nv.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_table_covers_the_header():
    bound = set(nv._SIGS) | {"hg_version", "hg_last_error", "hg_last_launch", "hg_launch_count", "hg_reset_launch_count", "hg_host_release"}
    assert set(declared_symbols()) <= bound, set(declared_symbols()) - bound


def test_version_and_error_string():
    L = nv.lib()
    assert L.hg_version() == 100
    # invalid arguments are rejected before any CUDA call: safe without a GPU
    rc = L.hg_rect2hex_bilinear(None, None, None, None, None, None, 1, 0, 4, 4, 4, nv.F32, nv.F32, 0, None)
    assert rc == -3 and b"bad shape" in L.hg_last_error()
    rc = L.hg_hexconv_out_shape(2, 2, 2, 1, 1, 0, C.byref(C.c_int64()), C.byref(C.c_int64()))
    assert rc == -3 and b"too small" in L.hg_last_error()
    ho, wo = C.c_int64(), C.c_int64()
    assert L.hg_hexconv_out_shape(256, 256, 2, 1, 1, 1, C.byref(ho), C.byref(wo)) == 0
    assert (ho.value, wo.value) == (256, 256)
    assert L.hg_hexconv_out_shape(8, 8, 3, 2, 1, 0, C.byref(ho), C.byref(wo)) == 0
    with pytest.raises(nv.HyGridNativeError):
        nv.call("hg_type_to_hex", None, None, 1, 4, 9, 3, nv.F32, nv.F32, None)


def test_argument_validation_without_a_gpu():
    """Every entry point validates its arguments before the first CUDA call and leaves a message in hg_last_error():
    error codes are part of the ABI (HG_E_ARG -1, HG_E_DTYPE -2, HG_E_SHAPE -3, HG_E_UNSUPPORTED -4), and empty
    inputs are a successful no-op."""
    L = nv.lib()
    err = lambda: L.hg_last_error().decode()
    # resampling: empty batch / empty output is fine, bad math mode and bad shapes are not
    assert L.hg_rect2hex_bilinear(None, None, None, None, None, None, 0, 8, 8, 8, 8, nv.F32, nv.F32, 0, None) == 0
    assert L.hg_hex2rect_linear(None, None, None, None, None, None, 3, 8, 8, 0, 8, nv.F32, nv.F32, 0, None) == 0
    assert L.hg_rect2hex_bilinear(None, None, None, None, None, None, 1, 8, 8, 8, 8, nv.F32, nv.F32, 7, None) == -1 and "math" in err()
    assert L.hg_hex2rect_nearest(None, None, None, None, 1, -2, 8, 8, 8, 4, None) == -3
    # pooling: unknown method, window leaving the image (the reference raises IndexError there)
    assert L.hg_hexpool_fwd(None, None, None, 1, 1, 8, 8, 4, 3, 2, 2, 2, 2, 2, 0, 0.0, 0, 0, 0.0, 9, nv.F32, None) == -1
    rc = L.hg_hexpool_fwd(None, None, None, 1, 1, 8, 8, 8, 7, 2, 2, 1, 1, 1, 0, 0.0, 0, 0, 0.0, nv.POOL_MAX, nv.F32, None)
    assert rc == -3 and "leaves the image" in err()
    assert L.hg_hexpool_fwd(None, None, None, 1, 0, 8, 8, 4, 3, 2, 2, 2, 2, 2, 0, 0.0, 0, 0, 0.0, nv.POOL_MAX, nv.F32, None) == 0
    # convolution descriptor checks
    bad = nv.ConvDesc(1, 4, 4, 8, 8, 9, 8, 2, 1, 1, 1, 1, 1, 0.0, nv.F32, nv.F32, 0, 0)  # wrong Ho
    assert L.hg_hexconv_dgrad(C.byref(bad), None, None, None, None) == -3 and "output shape" in err()
    bad = nv.ConvDesc(1, 4, 6, 8, 8, 8, 8, 2, 1, 1, 4, 1, 1, 0.0, nv.F32, nv.F32, 0, 0)  # out_channels % groups
    assert L.hg_hexconv_wgrad(C.byref(bad), None, None, None, None, None) == -1
    bad = nv.ConvDesc(1, 4, 4, 8, 8, 8, 8, 2, 1, 1, 1, 1, 1, 0.0, nv.F32, nv.F32, 2, 0)   # tcgen05 forced on 4 channels
    assert L.hg_hexconv_dgrad(C.byref(bad), None, None, None, None) == -4 and "tcgen05" in err()
    assert L.hg_hexconv_umma_eligible(C.byref(bad), 0) == 0
    # layout / padding / batch norm
    assert L.hg_pad2d(None, None, 1, 4, 4, 1, 1, 1, 1, 9, 0.0, nv.F32, None) == -1
    assert L.hg_pad2d(None, None, 1, 4, 4, 4, 0, 0, 0, 1, 0.0, nv.F32, None) == -3 and "reflect" in err()
    assert L.hg_bn_stats(None, None, 0, 4, 16, None) == -3
    assert L.hg_bn_apply(None, None, None, None, None, None, None, None, None, None, 2, 4, 16, C.c_float(1e-5), 0, None) == -1


def test_conv_out_shape_matches_oracle():
    from oracle import hexframes_oracle as HO
    L = nv.lib()
    for H in (5, 8, 9, 16, 33):
        for W in (5, 8, 13):
            for r in (2, 3):
                for s in (1, 2):
                    for d in (1, 2):
                        for pad in (0, 1, 2):
                            re_, ro, cols = HO.hexconv_out_shape(H + 2 * pad, W + 2 * pad, r, s, d)
                            ho, wo = C.c_int64(), C.c_int64()
                            rc = L.hg_hexconv_out_shape(H, W, r, s, d, pad, C.byref(ho), C.byref(wo))
                            ok = re_ > 0 and ro >= 0 and cols > 0 and re_ - ro in (0, 1)    # ro == 0: even rows only
                            assert (rc == 0) == ok
                            if ok:
                                assert (ho.value, wo.value) == (re_ + ro, cols)


def test_no_cpu_fallback():
    import torch
    from HyGrid import functional as Fn
    with pytest.raises(nv.HyGridNativeError):
        Fn.rect_to_hex(torch.zeros(1, 4, 4), None, "bilinear")


def test_ctypes_signatures_match_the_header_prototypes():
    """Arity and scalar/pointer kind of every ctypes signature agree with the prototype in include/hygrid_b200.h
    (a drifted table would pass garbage through the C ABI without any error)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = re.findall(r"\b(hg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(protos) >= 30
    scalars = {"int": C.c_int, "int64_t": C.c_int64, "double": C.c_double, "float": C.c_float}
    checked = 0
    for name, args in protos:
        if name not in nv._SIGS:
            continue
        kinds = []
        for a in args.split(","):
            a = " ".join(a.split())
            if a == "void":
                continue
            if "*" in a or "hg_stream_t" in a:
                kinds.append("ptr")
            else:
                kinds.append(scalars[a.replace("const ", "").rsplit(" ", 1)[0]])
        sig = nv._SIGS[name]
        assert len(sig) == len(kinds), (name, len(sig), len(kinds))
        for i, (s, k) in enumerate(zip(sig, kinds)):
            is_ptr = s is C.c_void_p or (isinstance(s, type) and issubclass(s, C._Pointer))
            assert (k == "ptr") == is_ptr and (k == "ptr" or s is k), (name, i, s, k)
        checked += 1
    assert checked == len(nv._SIGS)
