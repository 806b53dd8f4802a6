"""Drop-in surface of the torch layers and the numpy entry points that can be checked without a GPU: constructor
signatures, public attributes, state_dict keys, registry hooks and error types (reference
HexFrames.py:22-95, 255-283, HexModules.py:16-91, 97-253, geometry_np.py:12-25, 191-205, 358-371)."""
import numpy as np
import pytest
import torch

from HyGrid import HexFrames as hf
from HyGrid import HexModules as hm
from HyGrid import geometry_np as gnp
from HyGrid import geometry_torch as gt


def test_hexconv2d_attributes_and_parameters():
    m = hf.HexConv2d(6, 8, 1, 3, stride=2, padding=1, dilation=2, groups=2, bias=True)
    for name in ("in_channels", "out_channels", "even_odd_offset", "padded_even_odd_offset", "hexkernel_radius", "hexkernel_size",
                 "kernelnum", "stride", "sh", "sw", "out_even_odd_offset", "pad", "groups", "b", "dilation", "padding_mode",
                 "padding_value", "k_w", "k_h"):
        assert hasattr(m, name), name
    assert m.kernelnum == 3 * 3 * 3 - 3 * 3 + 1 == 19 and m.hexkernel_size == 5
    assert m.sw == 2 * m.stride and m.out_even_odd_offset == 0
    assert m.k_h == (m.hexkernel_size - 1) * 2 + 1 and m.k_w == 2 * 2 * (2 * 3 - 2) + 1
    assert m.padded_even_odd_offset == (1 + 1) % 2
    assert tuple(m.kernel.shape) == (8, 3, 1, 19) and tuple(m.bias.shape) == (8,)
    assert set(m.state_dict()) == {"kernel", "bias"}
    assert set(hf.HexConv2d(4, 4, 0, 2, bias=False).state_dict()) == {"kernel"}
    assert isinstance(m.extra_repr(), str)


def test_hexpool_attributes_and_errors():
    p = hf.HexPool2d("max", 2, 2)
    for name in ("out_offset", "offset", "method", "kernel_size", "kh", "kw", "stride", "sh", "sw", "padding", "padding_mode",
                 "padding_value", "ceil_mode", "count_include_pad"):
        assert hasattr(p, name), name
    assert (p.kh, p.kw, p.sh, p.sw) == (2, 2, 2, 2)
    q = hf.HexPool2d("average", (2, 3), (2, 4), padding=1, ceil_mode=True)
    assert (q.kh, q.kw, q.sh, q.sw) == (2, 3, 2, 4)
    assert hf.HexPool2d("min", 2).sh == 2                     # stride=None -> kernel_size (the reference crashes here)
    with pytest.raises(KeyError):
        hf.HexPool2d("median", 2, 2)
    with pytest.raises(Exception):
        hf.HexAdaptivePool2d([2, 2], "max")                   # only an int outsize is accepted (HexFrames.py:352-355)
    assert hf.HexGlobalPool2d("average") is not None and hf.HexAdaptivePool2d(3, "max") is not None


def test_builders_and_registry():
    assert "HexConv2d" in hm.CONV_LAYERS
    conv = hm.build_hexconv_layer(None, 4, 8, 0, 2)
    assert isinstance(conv, hf.HexConv2d) and conv.out_channels == 8
    conv = hm.build_hexconv_layer(dict(type="HexConv2d"), 4, 8, 1, 2, padding=1)
    assert conv.pad == 1 and conv.even_odd_offset == 1
    with pytest.raises(TypeError):
        hm.build_hexconv_layer("HexConv2d", 4, 8, 0, 2)
    with pytest.raises(KeyError):
        hm.build_hexconv_layer(dict(kind="HexConv2d"), 4, 8, 0, 2)
    with pytest.raises(KeyError):
        hm.build_hexconv_layer(dict(type="NoSuchConv"), 4, 8, 0, 2)
    name, norm = hm.build_hexnorm_layer(dict(type="BN"), 8)
    assert name == "bn" and isinstance(norm, torch.nn.BatchNorm2d) and norm.num_features == 8
    name, norm = hm.build_hexnorm_layer(dict(type="GN", num_groups=2), 8, postfix=1)
    assert name == "gn1" and isinstance(norm, torch.nn.GroupNorm)
    assert isinstance(hm.build_hexactivation_layer(dict(type="ReLU")), torch.nn.ReLU)
    assert hm.build_hexpadding_layer(dict(type="reflect"), 1) is not None


def test_hexconvmodule_structure():
    m = hm.HexConvModule(4, 8, 0, 2, padding=1, norm_cfg=dict(type="BN"))
    assert m.with_norm and m.with_activation and not m.with_bias          # bias='auto' -> not with_norm (HexModules.py:181-183)
    assert m.norm_name == "bn" and m.norm is m.bn and m.order == ("conv", "norm", "act")
    assert set(m.state_dict()) == {"conv.kernel", "bn.weight", "bn.bias", "bn.running_mean", "bn.running_var", "bn.num_batches_tracked"}
    m2 = hm.HexConvModule(4, 8, 0, 2, act_cfg=None, order=("act", "conv", "norm"))
    assert m2.with_bias and not m2.with_norm and not m2.with_activation
    assert float(m2.conv.bias.detach().abs().max()) == 0.0                          # init_weights zeroes the bias only
    for name in ("conv", "norm_name", "with_norm", "with_activation", "with_bias", "with_explicit_padding", "with_spectral_norm",
                 "order", "in_channels", "out_channels", "hexkernel_radius", "stride", "padding", "dilation", "groups"):
        assert hasattr(m, name), name
    with pytest.raises(AssertionError):
        hm.HexConvModule(4, 8, 0, 2, order=("conv", "norm"))


def test_numpy_entry_points_reject_bad_arguments_before_any_launch():
    img = np.zeros((3, 8, 8), np.float32)
    with pytest.raises(KeyError):
        gnp.rect_to_hex_resample(img, None, "cubic")
    with pytest.raises(KeyError):
        gnp.hex_to_rect_resample(img, None, "cubic")
    with pytest.raises(Exception):
        gnp.rect_to_hex_resample(np.zeros((2, 2, 2, 2)), None, "nearest")
    with pytest.raises(KeyError):
        gt.hex_to_square_resample(img, None, "cubic")
    if not torch.cuda.is_available():                          # the padding runs on the device: no silent CPU fallback
        with pytest.raises((RuntimeError, AssertionError)):
            gnp.heximpad(np.ones((5, 7, 3)), shape=(8, 9))


def test_out_buffer_is_validated_before_the_kernel_may_write_it():
    from HyGrid import functional as Fn
    like = torch.zeros(2, 3, 4, 5)
    assert Fn._result(None, (2, 3, 8, 9), torch.float64, like).shape == (2, 3, 8, 9)
    ok = torch.empty(2, 3, 8, 9)
    assert Fn._result(ok, (2, 3, 8, 9), None, like) is ok and Fn._result(ok, (2, 3, 8, 9), torch.float32, like) is ok
    with pytest.raises(ValueError):
        Fn._result(torch.empty(2, 3, 8, 8), (2, 3, 8, 9), None, like)
    with pytest.raises(ValueError):
        Fn._result(torch.empty(2, 3, 9, 8).transpose(2, 3), (2, 3, 8, 9), None, like)
    with pytest.raises(TypeError):
        Fn._result(ok, (2, 3, 8, 9), torch.float64, like)


def test_stream_argument_carries_its_device_through_ctypes():
    import ctypes as C
    from HyGrid import _native as nv
    st = nv.StreamArg(0x1234)
    st.device_index = 0
    assert isinstance(st, C.c_void_p) and st.value == 0x1234
    L = nv.lib()          # rejected for rows_step before the stream is touched: safe without a GPU
    assert L.hg_type_to_hex(None, None, 1, 4, 9, 3, nv.F32, nv.F32, st) == -1


def test_call_switches_to_the_stream_device_only_when_it_differs(monkeypatch):
    import contextlib
    from HyGrid import _native as nv
    entered = []

    @contextlib.contextmanager
    def fake_device(idx):
        entered.append(idx)
        yield

    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", fake_device)
    st = nv.StreamArg(0)
    for dev, want in ((0, []), (1, [1])):
        st.device_index = dev
        entered.clear()
        with pytest.raises(nv.HyGridNativeError):           # rows_step = 3 is rejected before anything touches the stream
            nv.call("hg_type_to_hex", None, None, 1, 4, 9, 3, nv.F32, nv.F32, st)
        assert entered == want
    entered.clear()
    with pytest.raises(nv.HyGridNativeError):               # host entry points end in an int, not a stream: never switch
        nv.call("hg_host_rect2hex", None, None, None, None, 1, 0, 4, 4, 4, nv.F32, nv.F32, 1, 0, 0)
    assert entered == []
