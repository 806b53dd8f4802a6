"""The row-streaming rect->hex kernel (csrc/hg_resample_stream.cu: same-pitch lattices, 16-byte loads / stores, neighbour
columns by warp shuffle) against the oracle: HG_MATH_FAST <= 1e-5 * max, HG_MATH_EXACT bit-identical (float64 result) or
its single rounding (float32 result), for shapes that exercise partial strips, several 1024-column CTA groups, band
boundaries, the row where i_n skips, and geometries the kernel must decline (they fall through to the tiled / direct
kernels and still match).  `hg_last_launch()` says which kernel took the call."""
import os

import numpy as np
import pytest
import torch

from oracle import hygrid_oracle as O

pytestmark = pytest.mark.gpu


def _run(Fn, nv, img, dsize, **kw):
    y = Fn.rect_to_hex(torch.from_numpy(img).cuda(), dsize, "bilinear", **kw)
    return y.cpu().numpy(), nv.last_launch()


@pytest.mark.parametrize("shape", [(2, 64, 128), (3, 37, 132), (1, 100, 1028), (2, 5, 4), (1, 130, 2052), (3, 257, 260),
                                   (1, 1030, 1024), (1, 2, 8), (2, 66, 4)])
def test_stream_kernel_same_size_vs_oracle(shape):
    from HyGrid import _native as nv
    from HyGrid import functional as Fn
    rng = np.random.default_rng(sum(shape))
    img = (rng.random(shape, dtype=np.float32) * 255).astype(np.float32)
    ref = np.stack([O.rect_to_hex_resample(img[i:i + 1], None, "bilinear") for i in range(shape[0])]).reshape(shape)
    for rows in ("8", "64"):
        os.environ["HG_R2H_STREAM_ROWS"] = rows
        fast, k = _run(Fn, nv, img, None, out_dtype=torch.float32, math="fast")
        assert k == "rect2hex_bilinear_stream", k
        assert float(np.abs(fast - ref).max()) <= 1e-5 * 255
        e64, k = _run(Fn, nv, img, None)                                   # float64, HG_MATH_EXACT: the drop-in call
        assert k == "rect2hex_bilinear_stream" and e64.dtype == np.float64 and np.array_equal(e64, ref)
        e32, k = _run(Fn, nv, img, None, out_dtype=torch.float32, math="exact")      # default: the TMA kernel keeps this variant
        assert np.array_equal(e32, ref.astype(np.float32))
        os.environ["HG_R2H_STREAM"] = "2"                                             # forced onto the streaming kernel
        try:
            e32, k = _run(Fn, nv, img, None, out_dtype=torch.float32, math="exact")
        finally:
            os.environ.pop("HG_R2H_STREAM", None)
        assert k == "rect2hex_bilinear_stream" and np.array_equal(e32, ref.astype(np.float32))
    os.environ.pop("HG_R2H_STREAM_ROWS", None)
    for pf in ("2", "4", "8"):                                                        # batch sizes of the batched-load kernel
        os.environ["HG_R2H_STREAM_PF"] = pf
        e64, k = _run(Fn, nv, img, None)
        assert k == "rect2hex_bilinear_stream" and np.array_equal(e64, ref)
        fast, k = _run(Fn, nv, img, None, out_dtype=torch.float32, math="fast")
        assert k == "rect2hex_bilinear_stream" and float(np.abs(fast - ref).max()) <= 1e-5 * 255
    os.environ.pop("HG_R2H_STREAM_PF", None)
    os.environ["HG_R2H_STREAM_V3"] = "0"                                              # the rolling-prefetch predecessor (A/B runs)
    try:
        e64, k = _run(Fn, nv, img, None)
        assert k == "rect2hex_bilinear_stream" and np.array_equal(e64, ref)
    finally:
        os.environ.pop("HG_R2H_STREAM_V3", None)
    # the same call with the streaming kernel switched off: tiled / direct kernels, identical exact result
    os.environ["HG_R2H_STREAM"] = "0"
    try:
        e64b, k = _run(Fn, nv, img, None)
        assert k != "rect2hex_bilinear_stream" and np.array_equal(e64b, ref)
    finally:
        os.environ.pop("HG_R2H_STREAM", None)


@pytest.mark.parametrize("case", [((2, 40, 64), (44, 64)), ((1, 64, 128), (60, 128)), ((1, 33, 256), (33, 252)),
                                  ((2, 48, 96), (24, 48)), ((1, 30, 100), (30, 100)), ((1, 64, 130), None)])
def test_other_geometries_decline_or_match(case):
    """Near-identity geometries run the streaming kernel when its column / row conditions hold; everything else (another
    pitch, widths that are not a multiple of four) is declined -- in both cases the result matches the oracle."""
    from HyGrid import _native as nv
    from HyGrid import functional as Fn
    shape, dsize = case
    rng = np.random.default_rng(7)
    img = (rng.random(shape, dtype=np.float32) * 255).astype(np.float32)
    ref = np.stack([np.asarray(O.rect_to_hex_resample(img[i:i + 1], dsize, "bilinear")).reshape((1,) + tuple(dsize or shape[1:]))
                    for i in range(shape[0])]).reshape((shape[0],) + tuple(dsize or shape[1:]))
    e64, _ = _run(Fn, nv, img, dsize)
    assert np.array_equal(e64, ref)
    fast, _ = _run(Fn, nv, img, dsize, out_dtype=torch.float32, math="fast")
    assert float(np.abs(fast - ref).max()) <= 1e-5 * 255


def test_stream_kernel_handles_non_finite_neighbours():
    """A NaN / inf cell must only reach the outputs whose taps touch it (the kernel selects taps, it never multiplies an
    unused neighbour by zero)."""
    from HyGrid import _native as nv
    from HyGrid import functional as Fn
    img = np.ones((1, 16, 128), np.float32)
    img[0, 5, 40] = np.nan
    img[0, 9, 77] = np.inf
    ref = O.rect_to_hex_resample(img, None, "bilinear").reshape(1, 16, 128)
    e64, k = _run(Fn, nv, img, None)
    assert k == "rect2hex_bilinear_stream"
    assert np.array_equal(np.isnan(e64), np.isnan(ref)) and np.array_equal(np.isinf(e64), np.isinf(ref))
    ok = np.isfinite(ref)
    assert np.array_equal(e64[ok], ref[ok])


# ----------------------------------------------------------------------------------------------------
# hex -> rect / hexresize between lattices of the same pitch (float32 weights)
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 64, 128), (3, 37, 132), (1, 100, 1028), (2, 5, 4), (1, 130, 2052), (3, 257, 260),
                                   (1, 1030, 1024), (1, 2, 8), (2, 66, 4)])
def test_hexsrc_stream_kernel_same_size_vs_oracle(shape):
    """HG_MATH_FAST hex->rect / hexresize with the output the size of the input: the row-streaming kernel evaluates the
    cell / triangle / weights of every sample in float32 from small quantities (|error| ~ 1e-7, continuous interpolant);
    the contract is 1e-5 of the range against the float64 oracle."""
    from HyGrid import _native as nv
    from HyGrid import functional as Fn
    rng = np.random.default_rng(sum(shape) + 1)
    img = (rng.random(shape, dtype=np.float32) * 255).astype(np.float32)
    x = torch.from_numpy(img).cuda()
    full = lambda fn: np.stack([np.asarray(fn(img[i:i + 1])).reshape(shape[1:]) for i in range(shape[0])])
    for twin in ("np", "torch"):
        ref = full(lambda a: O.hex_to_rect_resample(a, None, "linear", twin=twin))
        for pf in ("2", "4", "6"):
            os.environ["HG_H2R_STREAM_PF"] = pf
            got = Fn.hex_to_rect(x, None, "linear", out_dtype=torch.float32, math="fast", twin=twin)
            assert nv.last_launch() == "hexsrc_linear_stream", nv.last_launch()
            assert float(np.abs(got.cpu().numpy() - ref).max()) <= 1e-5 * 255
        os.environ.pop("HG_H2R_STREAM_PF", None)
    ref = full(lambda a: O.hexresize(a, shape[1:], "linear"))
    got = Fn.hex_resize(x, shape[1:], "linear", out_dtype=torch.float32, math="fast")
    assert nv.last_launch() == "hexsrc_linear_stream" and float(np.abs(got.cpu().numpy() - ref).max()) <= 1e-5 * 255
    # switched off: the TMA-tiled / direct kernels, same tolerance; exact mode never takes the streaming kernel
    os.environ["HG_H2R_STREAM"] = "0"
    try:
        got = Fn.hex_to_rect(x, None, "linear", out_dtype=torch.float32, math="fast", twin="np")
        assert nv.last_launch() != "hexsrc_linear_stream"
        assert float(np.abs(got.cpu().numpy() - full(lambda a: O.hex_to_rect_resample(a, None, "linear", twin="np"))).max()) <= 1e-5 * 255
    finally:
        os.environ.pop("HG_H2R_STREAM", None)
    e64 = Fn.hex_to_rect(x, None, "linear", twin="np")
    assert nv.last_launch() != "hexsrc_linear_stream"
    assert np.array_equal(e64.cpu().numpy(), full(lambda a: O.hex_to_rect_resample(a, None, "linear", twin="np")))


@pytest.mark.parametrize("case", [((2, 40, 64), (44, 64)), ((1, 64, 128), (128, 256)), ((1, 33, 256), (33, 252)), ((2, 48, 96), (24, 48))])
def test_hexsrc_other_geometries_decline_and_match(case):
    from HyGrid import _native as nv
    from HyGrid import functional as Fn
    shape, dsize = case
    rng = np.random.default_rng(11)
    img = (rng.random(shape, dtype=np.float32) * 255).astype(np.float32)
    ref = np.stack([np.asarray(O.hex_to_rect_resample(img[i:i + 1], dsize, "linear", twin="np")).reshape(dsize) for i in range(shape[0])])
    got = Fn.hex_to_rect(torch.from_numpy(img).cuda(), dsize, "linear", out_dtype=torch.float32, math="fast", twin="np")
    assert float(np.abs(got.cpu().numpy() - ref).max()) <= 1e-5 * 255
