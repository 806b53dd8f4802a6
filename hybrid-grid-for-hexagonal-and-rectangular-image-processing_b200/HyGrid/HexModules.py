"""``HyGrid.HexModules`` (/root/reference/HyGrid/HexModules.py): mmcv-style conv / norm / activation bundle and
the ``CONV_LAYERS`` registry hook, over the sm_100a hex convolution.  Same builders, ``HexConvModule``
constructor, attributes, ``order`` handling and error types.  mmcv 1.x is used when importable; otherwise
the in-repo ``_registry`` shim supplies the six helpers the reference imports (HexModules.py:7-12).

Explicit padding layers ('zero' / 'reflect' / 'replicate') are the pad kernel of libhygrid_b200.so.  In
inference (no grad) a ``conv -> ReLU`` pair without a norm in between runs as one launch (fused epilogue)."""
from __future__ import annotations

import warnings
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import HexFrames as hnn
from . import _norm

try:  # pragma: no cover - mmcv 1.x is not installed in the build image
    from mmcv.cnn.bricks.norm import build_norm_layer
    from mmcv.cnn.bricks.padding import build_padding_layer
    from mmcv.cnn.bricks.activation import build_activation_layer
    from mmcv.cnn.bricks.registry import CONV_LAYERS, PADDING_LAYERS
    from mmcv.utils import _BatchNorm, _InstanceNorm
    from mmcv.cnn.utils import constant_init, kaiming_init
    _HAVE_MMCV = True
except Exception:
    from ._registry import (CONV_LAYERS, PADDING_LAYERS, _BatchNorm, _InstanceNorm, build_activation_layer,
                            build_norm_layer, build_padding_layer, constant_init, kaiming_init)
    _HAVE_MMCV = False

__all__ = ["build_hexconv_layer", "build_hexpadding_layer", "build_hexnorm_layer", "build_hexactivation_layer",
           "HexConvModule", "CONV_LAYERS", "PADDING_LAYERS"]


class _HexPad2d(nn.Module):
    """Padding layer on the pad kernel (same constructor as nn.ZeroPad2d / ReflectionPad2d / ReplicationPad2d)."""
    _mode = "constant"

    def __init__(self, padding):
        super().__init__()
        self.padding = (padding,) * 4 if isinstance(padding, int) else tuple(padding)

    def forward(self, x):
        pl, pr, pt, pb = self.padding
        return hnn._pad4(x, pl, pr, pt, pb, self._mode, 0)

    def extra_repr(self):
        return f"{self.padding}"


class HexZeroPad2d(_HexPad2d):
    _mode = "constant"


class HexReflectionPad2d(_HexPad2d):
    _mode = "reflect"


class HexReplicationPad2d(_HexPad2d):
    _mode = "replicate"


if 'HexConv2d' not in CONV_LAYERS:
    CONV_LAYERS.register_module('HexConv2d', module=hnn.HexConv2d)
if 'HexConv2dAdaptivePadding' not in CONV_LAYERS:
    CONV_LAYERS.register_module('HexConv2dAdaptivePadding', module=hnn.HexConv2dAdaptivePadding)
if not _HAVE_MMCV:
    for _n, _m in (('zero', HexZeroPad2d), ('reflect', HexReflectionPad2d), ('replicate', HexReplicationPad2d)):
        if _n not in PADDING_LAYERS:
            PADDING_LAYERS.register_module(_n, module=_m)


def build_hexconv_layer(cfg: Optional[Dict], *args, **kwargs) -> nn.Module:
    """HexModules.py:22-54."""
    if cfg is None:
        cfg_ = dict(type='HexConv2d')
    else:
        if not isinstance(cfg, dict):
            raise TypeError('cfg must be a dict')
        if 'type' not in cfg:
            raise KeyError('the cfg dict must contain the key "type"')
        cfg_ = cfg.copy()
    layer_type = cfg_.pop('type')
    if layer_type not in CONV_LAYERS:
        raise KeyError(f'Unrecognized layer type {layer_type}')
    conv_layer = CONV_LAYERS.get(layer_type)
    return conv_layer(*args, **kwargs, **cfg_)


def build_hexpadding_layer(cfg: Dict, *args, **kwargs) -> nn.Module:
    """HexModules.py:56-67."""
    return build_padding_layer(cfg, *args, **kwargs)


def build_hexnorm_layer(cfg: Dict, num_features: int, postfix: Union[int, str] = '') -> Tuple[str, nn.Module]:
    """HexModules.py:69-89."""
    return build_norm_layer(cfg, num_features, postfix)


def build_hexactivation_layer(cfg: Dict) -> nn.Module:
    """HexModules.py:90-91."""
    return build_activation_layer(cfg)


class HexConvModule(nn.Module):
    """A hexconv block that bundles hexconv / norm / activation layers (HexModules.py:97-288)."""

    _abbr_ = 'conv_block'

    def __init__(self, in_channels: int, out_channels: int, even_odd_offset: int, hexkernel_radius: int,
                 stride: int = 1, padding: int = 0, dilation: int = 1, groups: int = 1,
                 bias: Union[bool, str] = 'auto', conv_cfg: Optional[Dict] = None, norm_cfg: Optional[Dict] = None,
                 act_cfg: Optional[Dict] = dict(type='ReLU'), inplace: bool = True, with_spectral_norm: bool = False,
                 padding_mode: str = 'zeros', order: tuple = ('conv', 'norm', 'act')):
        super().__init__()
        assert conv_cfg is None or isinstance(conv_cfg, dict)
        assert norm_cfg is None or isinstance(norm_cfg, dict)
        assert act_cfg is None or isinstance(act_cfg, dict)
        official_padding_mode = ['zeros', 'circular']
        self.conv_cfg = conv_cfg
        self.norm_cfg = norm_cfg
        self.act_cfg = act_cfg
        self.inplace = inplace
        self.with_spectral_norm = with_spectral_norm
        self.with_explicit_padding = padding_mode not in official_padding_mode
        self.order = order
        assert isinstance(self.order, tuple) and len(self.order) == 3
        assert set(order) == {'conv', 'norm', 'act'}

        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.with_bias = bias

        if self.with_explicit_padding:
            pad_cfg = dict(type=padding_mode)
            self.padding_layer = build_hexpadding_layer(pad_cfg, padding)

        conv_padding = 0 if self.with_explicit_padding else padding
        self.conv = build_hexconv_layer(conv_cfg, in_channels, out_channels, even_odd_offset, hexkernel_radius,
                                        stride=stride, padding=conv_padding, dilation=dilation, groups=groups, bias=bias)
        self.in_channels = self.conv.in_channels
        self.out_channels = self.conv.out_channels
        self.hexkernel_radius = self.conv.hexkernel_radius
        self.stride = self.conv.stride
        self.padding = padding
        self.dilation = self.conv.dilation
        self.groups = self.conv.groups

        if self.with_spectral_norm:
            # the reference calls nn.utils.spectral_norm(self.conv), which looks for a parameter called
            # "weight" and fails on HexConv2d's "kernel" (HexModules.py:204-205); name it explicitly.
            self.conv = nn.utils.spectral_norm(self.conv, name='kernel')

        if self.with_norm:
            if order.index('norm') > order.index('conv'):
                norm_channels = out_channels
            else:
                norm_channels = in_channels
            self.norm_name, norm = build_hexnorm_layer(norm_cfg, norm_channels)
            self.add_module(self.norm_name, norm)
            if self.with_bias:
                if isinstance(norm, (_BatchNorm, _InstanceNorm)):
                    warnings.warn('Unnecessary conv bias before batch/instance norm')
        else:
            self.norm_name = None

        if self.with_activation:
            act_cfg_ = act_cfg.copy()
            if act_cfg_['type'] not in ['Tanh', 'PReLU', 'Sigmoid', 'HSigmoid', 'Swish', 'GELU']:
                act_cfg_.setdefault('inplace', inplace)
            self.activate = build_hexactivation_layer(act_cfg_)

        self.init_weights()

    @property
    def norm(self):
        if self.norm_name:
            return getattr(self, self.norm_name)
        return None

    def init_weights(self):
        """HexModules.py:254-273: mmcv's kaiming_init only touches ``weight`` / ``bias``; HexConv2d's parameter is
        called ``kernel``, so effectively the conv keeps its own kaiming-uniform kernel and the bias is zeroed."""
        if not hasattr(self.conv, 'init_weights'):
            if self.with_activation and self.act_cfg['type'] == 'LeakyReLU':
                nonlinearity = 'leaky_relu'
                a = self.act_cfg.get('negative_slope', 0.01)
            else:
                nonlinearity = 'relu'
                a = 0
            kaiming_init(self.conv, a=a, nonlinearity=nonlinearity)
        if self.with_norm:
            constant_init(self.norm, 1, bias=0)

    def _fusable_relu(self, idx, x, activate, norm):
        """conv immediately followed by a plain ReLU, nothing recorded for backward."""
        if torch.is_grad_enabled() or self.with_spectral_norm or not isinstance(self.conv, hnn.HexConv2d):
            return False
        if not (activate and self.with_activation and type(self.activate) is nn.ReLU):
            return False
        rest = [l for l in self.order[idx + 1:] if not (l == 'norm' and not (norm and self.with_norm))]
        return bool(rest) and rest[0] == 'act'

    def _fusable_bn(self, idx, activate, norm):
        """Inference: conv -> BatchNorm(eval, running statistics) [-> ReLU] collapses into the conv epilogue
        (hg_hexconv_fwd_affine): returns (scale, shift, relu) or None.  ref: HexModules.py:275-288."""
        if torch.is_grad_enabled() or self.with_spectral_norm or not isinstance(self.conv, hnn.HexConv2d):
            return None
        if not (norm and self.with_norm) or self.order[idx + 1:idx + 2] != ('norm',):
            return None
        bn = self.norm
        if not isinstance(bn, nn.modules.batchnorm._BatchNorm) or bn.training or bn.running_mean is None:
            return None
        scale = torch.rsqrt(bn.running_var.float() + bn.eps)
        if bn.weight is not None:
            scale = scale * bn.weight.float()
        shift = -bn.running_mean.float() * scale
        if bn.bias is not None:
            shift = shift + bn.bias.float()
        if self.conv.bias is not None:
            shift = shift + self.conv.bias.float() * scale
        tail = self.order[idx + 2:]
        relu = bool(tail) and tail[0] == 'act' and activate and self.with_activation and type(self.activate) is nn.ReLU
        return scale, shift, relu

    def _frame(self):
        """The explicit padding layer as a frame the conv kernel resolves itself (no padded copy): ``(p, mode)`` when the
        layer is one of this module's pad layers with the same padding on all four sides in front of a plain HexConv2d,
        else None.  ref: HexModules.py:185-190, :279-281."""
        if not self.with_explicit_padding or self.with_spectral_norm or type(self.conv) is not hnn.HexConv2d or self.conv.pad:
            return None
        layer = self.padding_layer
        if not isinstance(layer, _HexPad2d) or len(set(layer.padding)) != 1:
            return None
        return (int(layer.padding[0]), layer._mode) if layer.padding[0] > 0 else None

    def forward(self, x: torch.Tensor, activate: bool = True, norm: bool = True) -> torch.Tensor:
        fused = False
        skip_norm = False
        for idx, layer in enumerate(self.order):
            if layer == 'conv':
                kw = {}
                if self.with_explicit_padding:
                    frame = self._frame()
                    if frame is not None and x.dim() >= 3 and (frame[1] != 'reflect' or frame[0] < min(x.shape[-2:])):
                        kw = dict(frame=frame)
                    else:
                        x = self.padding_layer(x)
                bn = self._fusable_bn(idx, activate, norm)
                if bn is not None:
                    scale, shift, fused = bn
                    skip_norm = True
                    x = self.conv(x, relu=fused, affine=(scale, shift), **kw)
                    continue
                fused = self._fusable_relu(idx, x, activate, norm)
                x = self.conv(x, relu=True, **kw) if fused else self.conv(x, **kw)
            elif layer == 'norm' and norm and self.with_norm:
                if skip_norm:
                    skip_norm = False
                    continue
                if _norm.bn_supported(self.norm, x):
                    # BatchNorm2d (+ a directly following plain ReLU) on the library's streaming kernels: two passes
                    # forward, two backward, instead of cuDNN batch-norm + ReLU and their backward twins
                    tail = self.order[idx + 1:]
                    fused = bool(tail) and tail[0] == 'act' and activate and self.with_activation and type(self.activate) is nn.ReLU
                    x = _norm.batch_norm_relu(self.norm, x, relu=fused)
                    continue
                x = self.norm(x)
            elif layer == 'act' and activate and self.with_activation:
                if fused:
                    fused = False
                    continue
                x = self.activate(x)
        return x
