"""HAHI neck in the configuration the DDIM heads ship (cross_att=False, self_att=False): a purely
feed-forward conv pyramid (reference src/model/necks/hahi.py:165-276).  The reference also computes sine
position encodings, reference points and masks on every forward and discards them, and builds two
MultiScaleDeformableAttention modules it never calls; the mirror keeps their *parameters* (so reference
checkpoints load key-for-key) and skips the dead work.  Step-invariant: runs once per image."""
import math

import torch
import torch.nn as nn

from .._blocks import ConvModule


class _DeformAttnParams(nn.Module):
    """Parameter container with mmcv MultiScaleDeformableAttention's names (hahi.py:109-118)."""

    def __init__(self, dim, levels, heads, points):
        super().__init__()
        self.sampling_offsets = nn.Linear(dim, heads * levels * points * 2)
        self.attention_weights = nn.Linear(dim, heads * levels * points)
        self.value_proj = nn.Linear(dim, dim)
        self.output_proj = nn.Linear(dim, dim)
        nn.init.zeros_(self.sampling_offsets.weight)
        theta = torch.arange(heads, dtype=torch.float32) * (2.0 * math.pi / heads)
        grid = torch.stack([theta.cos(), theta.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(heads, 1, 1, 2).repeat(1, levels, points, 1)
        grid = grid * torch.arange(1, points + 1, dtype=torch.float32).view(1, 1, points, 1)
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.reshape(-1))
        nn.init.zeros_(self.attention_weights.weight)
        nn.init.zeros_(self.attention_weights.bias)
        for lin in (self.value_proj, self.output_proj):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)


class HAHIHeteroNeck(nn.Module):
    def __init__(self, in_channels, out_channels, embedding_dim, positional_encoding=None, scales=(1, 1, 1, 1),
                 cross_att=False, self_att=False, num_points=8, **unused):
        super().__init__()
        if cross_att or self_att:
            raise NotImplementedError("the shipped DiffusionDepth heads disable both deformable attentions "
                                      "(head ctor :54-56); only that configuration is served")
        if any(s != 1 for s in scales):
            raise NotImplementedError("scales != 1 are not used by the shipped heads")
        self.in_channels, self.out_channels, self.embedding_dim = list(in_channels), list(out_channels), embedding_dim
        e = embedding_dim
        self.lateral_convs = nn.ModuleList(ConvModule(i, o, 1) for i, o in zip(in_channels, out_channels))
        self.trans_proj = nn.ModuleList(ConvModule(o, e, 1) for o in out_channels[1:])
        self.trans_fusion = nn.ModuleList(ConvModule(o + e, o, 3, padding=1) for o in out_channels[1:])
        self.conv_proj = nn.Sequential(ConvModule(in_channels[0], e, 1))
        self.conv_fusion = nn.Sequential(ConvModule(in_channels[0] + e, out_channels[0], 3, padding=1))
        # present in the reference state_dict, never used by its forward:
        self.reference_points = nn.Linear(e, 2)
        self.level_embed = nn.Parameter(torch.zeros(4, e))  # reference leaves this uninitialised (hahi.py:107)
        self.multi_att = _DeformAttnParams(e, 4, 8, num_points)
        self.self_attn = _DeformAttnParams(e, 4, 8, num_points)

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        lat = [conv(x) for conv, x in zip(self.lateral_convs, inputs)]
        # level 0 ("conv" branch): concat order is [projection, lateral]  (hahi.py:226-250)
        outs = [self.conv_fusion(torch.cat([self.conv_proj(lat[0]), lat[0]], dim=1))]
        # levels 1..3 ("transformer" branch): concat order is [lateral, projection]  (:253-272)
        for i, f in enumerate(lat[1:]):
            outs.append(self.trans_fusion[i](torch.cat([f, self.trans_proj[i](f)], dim=1)))
        return tuple(outs)
