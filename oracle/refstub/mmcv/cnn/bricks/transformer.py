"""FFN / build_dropout / positional-encoding builder (mmcv 1.x + mmdet SinePositionalEncoding)."""
import math
import torch
import torch.nn as nn
from mmcv.runner.base_module import BaseModule, Sequential
from mmcv.utils import Registry

POSITIONAL_ENCODING = Registry('position encoding')


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.1):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x.div(keep) * mask


def build_dropout(cfg, default_args=None):
    if cfg is None:
        return nn.Identity()
    cfg = dict(cfg)
    typ = cfg.pop('type')
    if typ == 'DropPath':
        return DropPath(**cfg)
    if typ == 'Dropout':
        return nn.Dropout(cfg.get('drop_prob', cfg.get('p', 0.5)))
    raise KeyError(typ)


class FFN(BaseModule):
    """Linear -> act -> drop -> Linear -> drop, + identity.  Keys: layers.0.0.*, layers.1.*"""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        from mmcv.cnn import build_activation_layer
        self.activate = build_activation_layer(act_cfg)
        layers, c = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(Sequential(nn.Linear(c, feedforward_channels), self.activate,
                                     nn.Dropout(ffn_drop)))
            c = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = Sequential(*layers)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


@POSITIONAL_ENCODING.register_module()
class SinePositionalEncoding(BaseModule):
    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * math.pi,
                 eps=1e-6, offset=0., init_cfg=None):
        super().__init__(init_cfg)
        self.num_feats, self.temperature = num_feats, temperature
        self.normalize, self.scale, self.eps, self.offset = normalize, scale, eps, offset

    def forward(self, mask):
        mask = mask.to(torch.int)
        not_mask = 1 - mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            y_embed = (y_embed + self.offset) / (y_embed[:, -1:, :] + self.eps) * self.scale
            x_embed = (x_embed + self.offset) / (x_embed[:, :, -1:] + self.eps) * self.scale
        dim_t = torch.arange(self.num_feats, dtype=torch.float32, device=mask.device)
        dim_t = self.temperature ** (2 * (dim_t // 2) / self.num_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        B, H, W = mask.size()
        pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).view(B, H, W, -1)
        pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).view(B, H, W, -1)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def build_positional_encoding(cfg, default_args=None):
    return POSITIONAL_ENCODING.build(cfg)
