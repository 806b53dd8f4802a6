"""Minimal stand-in for the six mmcv-1.x helpers ``HexModules`` imports (HexModules.py:7-12) -- mmcv is
not installable here.  Same call signatures and naming rules (``build_norm_layer`` returns
``(abbr + postfix, layer)``; ``kaiming_init`` only touches ``module.weight`` / ``module.bias``).
If a real mmcv 1.x is importable, HexModules uses it instead of this file."""
from __future__ import annotations

import inspect

import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm
from torch.nn.modules.instancenorm import _InstanceNorm

__all__ = ["Registry", "CONV_LAYERS", "PADDING_LAYERS", "NORM_LAYERS", "ACTIVATION_LAYERS", "build_norm_layer",
           "build_padding_layer", "build_activation_layer", "constant_init", "kaiming_init", "_BatchNorm", "_InstanceNorm"]


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    def __contains__(self, key):
        return key in self._module_dict

    def __len__(self):
        return len(self._module_dict)

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        if not isinstance(force, bool):
            raise TypeError(f'force must be a boolean, but got {type(force)}')

        def _register(cls):
            names = [name] if isinstance(name, str) else (name or [cls.__name__])
            for n in names:
                if not force and n in self._module_dict:
                    raise KeyError(f'{n} is already registered in {self._name}')
                self._module_dict[n] = cls
            return cls
        if module is not None:
            _register(module)
            return module
        return _register


CONV_LAYERS = Registry('conv layer')
PADDING_LAYERS = Registry('padding layer')
NORM_LAYERS = Registry('norm layer')
ACTIVATION_LAYERS = Registry('activation layer')

for _n, _m in (('Conv1d', nn.Conv1d), ('Conv2d', nn.Conv2d), ('Conv3d', nn.Conv3d), ('Conv', nn.Conv2d)):
    CONV_LAYERS.register_module(_n, module=_m)
for _n, _m in (('BN', nn.BatchNorm2d), ('BN1d', nn.BatchNorm1d), ('BN2d', nn.BatchNorm2d), ('BN3d', nn.BatchNorm3d),
               ('SyncBN', nn.SyncBatchNorm), ('GN', nn.GroupNorm), ('LN', nn.LayerNorm), ('IN', nn.InstanceNorm2d),
               ('IN1d', nn.InstanceNorm1d), ('IN2d', nn.InstanceNorm2d), ('IN3d', nn.InstanceNorm3d)):
    NORM_LAYERS.register_module(_n, module=_m)
for _m in (nn.ReLU, nn.LeakyReLU, nn.PReLU, nn.RReLU, nn.ReLU6, nn.ELU, nn.Sigmoid, nn.Tanh, nn.GELU):
    ACTIVATION_LAYERS.register_module(module=_m)
ACTIVATION_LAYERS.register_module('Swish', module=nn.SiLU)
ACTIVATION_LAYERS.register_module('HSigmoid', module=nn.Hardsigmoid)
ACTIVATION_LAYERS.register_module('HSwish', module=nn.Hardswish)


def _infer_abbr(class_type):
    if not inspect.isclass(class_type):
        raise TypeError(f'class_type must be a type, but got {type(class_type)}')
    if hasattr(class_type, '_abbr_'):
        return class_type._abbr_
    if issubclass(class_type, _InstanceNorm):
        return 'in'
    if issubclass(class_type, _BatchNorm):
        return 'bn'
    if issubclass(class_type, nn.GroupNorm):
        return 'gn'
    if issubclass(class_type, nn.LayerNorm):
        return 'ln'
    n = class_type.__name__.lower()
    for k in ('batch', 'group', 'layer', 'instance'):
        if k in n:
            return {'batch': 'bn', 'group': 'gn', 'layer': 'ln', 'instance': 'in'}[k]
    return 'norm_layer'


def build_norm_layer(cfg, num_features, postfix=''):
    if not isinstance(cfg, dict):
        raise TypeError('cfg must be a dict')
    if 'type' not in cfg:
        raise KeyError('the cfg dict must contain the key "type"')
    cfg_ = cfg.copy()
    layer_type = cfg_.pop('type')
    if layer_type not in NORM_LAYERS:
        raise KeyError(f'Unrecognized norm type {layer_type}')
    norm_layer = NORM_LAYERS.get(layer_type)
    abbr = _infer_abbr(norm_layer)
    assert isinstance(postfix, (int, str))
    name = abbr + str(postfix)
    requires_grad = cfg_.pop('requires_grad', True)
    cfg_.setdefault('eps', 1e-5)
    if layer_type != 'GN':
        layer = norm_layer(num_features, **cfg_)
        if layer_type == 'SyncBN' and hasattr(layer, '_specify_ddp_gpu_num'):
            layer._specify_ddp_gpu_num(1)
    else:
        assert 'num_groups' in cfg_
        layer = norm_layer(num_channels=num_features, **cfg_)
    for param in layer.parameters():
        param.requires_grad = requires_grad
    return name, layer


def build_padding_layer(cfg, *args, **kwargs):
    if not isinstance(cfg, dict):
        raise TypeError('cfg must be a dict')
    if 'type' not in cfg:
        raise KeyError('the cfg dict must contain the key "type"')
    cfg_ = cfg.copy()
    padding_type = cfg_.pop('type')
    if padding_type not in PADDING_LAYERS:
        raise KeyError(f'Unrecognized padding type {padding_type}.')
    return PADDING_LAYERS.get(padding_type)(*args, **kwargs, **cfg_)


def build_activation_layer(cfg):
    if not isinstance(cfg, dict):
        raise TypeError('cfg must be a dict')
    if 'type' not in cfg:
        raise KeyError('the cfg dict must contain the key "type"')
    cfg_ = cfg.copy()
    act_type = cfg_.pop('type')
    if act_type not in ACTIVATION_LAYERS:
        raise KeyError(f'Unrecognized activation type {act_type}')
    return ACTIVATION_LAYERS.get(act_type)(**cfg_)


def constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode='fan_out', nonlinearity='relu', bias=0, distribution='normal'):
    assert distribution in ['uniform', 'normal']
    if hasattr(module, 'weight') and module.weight is not None:
        if distribution == 'uniform':
            nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        else:
            nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)
