"""mmcv.cnn stand-ins: builders + ConvModule + weight-init helpers (mmcv-full 1.x semantics)."""
import torch.nn as nn
from .utils.weight_init import (constant_init, kaiming_init, xavier_init,  # noqa: F401
                                trunc_normal_init, normal_init)

_CONV = {None: nn.Conv2d, 'Conv2d': nn.Conv2d, 'Conv': nn.Conv2d, 'Conv1d': nn.Conv1d}
_ACT = {'ReLU': nn.ReLU, 'LeakyReLU': nn.LeakyReLU, 'GELU': nn.GELU, 'Sigmoid': nn.Sigmoid,
        'Tanh': nn.Tanh}


def build_conv_layer(cfg, *args, **kwargs):
    typ = None if cfg is None else dict(cfg).get('type')
    return _CONV[typ](*args, **kwargs)


def build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    requires_grad = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    if typ in ('BN', 'BN2d', 'SyncBN'):
        layer, abbr = nn.BatchNorm2d(num_features, **cfg), 'bn'
    elif typ == 'LN':
        layer, abbr = nn.LayerNorm(num_features, **cfg), 'ln'
    elif typ == 'GN':
        layer, abbr = nn.GroupNorm(num_channels=num_features, **cfg), 'gn'
    else:
        raise KeyError(typ)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer


def build_upsample_layer(cfg, *args, **kwargs):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    if typ == 'deconv':
        return nn.ConvTranspose2d(*args, **cfg, **kwargs)
    if typ in ('nearest', 'bilinear'):
        return nn.Upsample(*args, mode=typ, **cfg, **kwargs)
    raise KeyError(typ)


def build_activation_layer(cfg):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    if typ == 'GELU':
        cfg.pop('inplace', None)
    return _ACT[typ](**cfg)


def build_plugin_layer(cfg, postfix='', **kwargs):
    raise NotImplementedError('plugins are not used by the reference hot path')


class ConvModule(nn.Module):
    """conv -> norm -> act; bias='auto' means bias iff no norm; kaiming(fan_out) init."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias='auto', conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'),
                 inplace=True, with_spectral_norm=False, padding_mode='zeros',
                 order=('conv', 'norm', 'act')):
        super().__init__()
        assert order == ('conv', 'norm', 'act') and padding_mode == 'zeros'
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.with_bias = bias
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size,
                                     stride=stride, padding=padding, dilation=dilation,
                                     groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        else:
            self.norm_name = None
        if self.with_activation:
            a = dict(act_cfg)
            if a['type'] not in ('Tanh', 'PReLU', 'Sigmoid', 'HSigmoid', 'Swish', 'GELU'):
                a.setdefault('inplace', inplace)
            self.activate = build_activation_layer(a)
        self.init_weights()

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.norm_name else None

    def init_weights(self):
        if self.with_activation and self.act_cfg['type'] == 'LeakyReLU':
            nonlinearity, a = 'leaky_relu', self.act_cfg.get('negative_slope', 0.01)
        else:
            nonlinearity, a = 'relu', 0
        kaiming_init(self.conv, a=a, nonlinearity=nonlinearity)
        if self.with_norm:
            constant_init(self.norm, 1, bias=0)

    def forward(self, x, activate=True, norm=True):
        x = self.conv(x)
        if norm and self.with_norm:
            x = self.norm(x)
        if activate and self.with_activation:
            x = self.activate(x)
        return x
