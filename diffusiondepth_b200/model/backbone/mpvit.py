"""MPViT (multi-path vision transformer) backbone with the reference's parameter layout (reference
src/model/backbone/mpvit.py:57-741; the variant the `DDIMDepthEstimate_MPVIT_ADDHAHI` head is trained with is
`mpvit_small`, README.md:272).  Host-side mirror: a torch module whose `state_dict` loads the released checkpoints
key for key (including the duplicated keys of the position-encoding modules each encoder shares with its blocks).
On a CUDA input the plugin does not call this module's forward: the engine runs the network itself from these
parameters (`dd_run_backbone`, kind DD_BACKBONE_MPVIT — csrc/mpvit.cuh + the GEMM path); the torch forward
below (under `exact_fp32()`) is what `native_backbone = False` and shapes the engine does not instantiate fall back to.

Structure per stage i (dims C_i -> C_{i+1}):
  patch_embed_stages[i]: a CHAIN of depthwise-separable 3x3 embeddings (dw 3x3, pw 1x1, BN, Hardswish); the first
      has stride 2, embedding p consumes the output of embedding p-1 and each output feeds one path (:212-238);
  mhca_stages[i]: one conv path `InvRes` on the first embedding + `num_path` transformer encoders (factorised
      attention with convolutional relative position encoding), concatenated and fused by a 1x1 `aggregate`
      (:535-583).
The stem keeps full resolution (both strides are 1 in the reference, :627-643), so the four outputs sit at 1/2, 1/4,
1/8 and 1/16 of the image."""
import math
import os

import torch
import torch.nn as nn

from .._blocks import DropPath


def _fan_out_normal_(conv):
    fan_out = conv.kernel_size[0] * conv.kernel_size[1] * conv.out_channels
    nn.init.normal_(conv.weight, 0.0, math.sqrt(2.0 / fan_out))


class Conv2d_BN(nn.Module):
    """conv (no bias) -> BatchNorm2d -> optional activation; keys `conv.*`, `bn.*` (:85-122)."""

    def __init__(self, in_ch, out_ch, kernel_size=1, stride=1, pad=0, act_layer=None):
        super().__init__()
        self.conv = nn.Conv2d(in_ch, out_ch, kernel_size, stride, pad, bias=False)
        self.bn = nn.BatchNorm2d(out_ch)
        _fan_out_normal_(self.conv)
        self.act_layer = act_layer() if act_layer is not None else nn.Identity()

    def forward(self, x):
        return self.act_layer(self.bn(self.conv(x)))


class DWConv2d_BN(nn.Module):
    """depthwise k x k -> pointwise 1x1 -> BN -> Hardswish; keys `dwconv`, `pwconv`, `bn` (:125-175)."""

    def __init__(self, in_ch, out_ch, kernel_size=1, stride=1):
        super().__init__()
        self.dwconv = nn.Conv2d(in_ch, out_ch, kernel_size, stride, (kernel_size - 1) // 2, groups=out_ch, bias=False)
        self.pwconv = nn.Conv2d(out_ch, out_ch, 1, 1, 0, bias=False)
        self.bn = nn.BatchNorm2d(out_ch)
        self.act = nn.Hardswish()
        _fan_out_normal_(self.dwconv)
        _fan_out_normal_(self.pwconv)

    def forward(self, x):
        return self.act(self.bn(self.pwconv(self.dwconv(x))))


class DWCPatchEmbed(nn.Module):
    def __init__(self, in_chans, embed_dim, patch_size, stride):
        super().__init__()
        self.patch_conv = DWConv2d_BN(in_chans, embed_dim, kernel_size=patch_size, stride=stride)

    def forward(self, x):
        return self.patch_conv(x)


class Patch_Embed_stage(nn.Module):
    def __init__(self, embed_dim, num_path=4, isPool=False):
        super().__init__()
        self.patch_embeds = nn.ModuleList(
            DWCPatchEmbed(embed_dim, embed_dim, patch_size=3, stride=2 if isPool and idx == 0 else 1)
            for idx in range(num_path))

    def forward(self, x):
        outs = []
        for pe in self.patch_embeds:  # a chain, not parallel branches
            x = pe(x)
            outs.append(x)
        return outs


class ConvPosEnc(nn.Module):
    """x + depthwise3x3(x) on the token map (:241-259)."""

    def __init__(self, dim, k=3):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, k, 1, k // 2, groups=dim)

    def forward(self, x, size):
        B, N, C = x.shape
        feat = x.transpose(1, 2).reshape(B, C, *size)
        return (self.proj(feat) + feat).flatten(2).transpose(1, 2)


class ConvRelPosEnc(nn.Module):
    """q * depthwise_conv(v): the heads are split into groups, each with its own window (:262-330)."""

    def __init__(self, Ch, h, window):
        super().__init__()
        window = {window: h} if isinstance(window, int) else dict(window)
        self.conv_list = nn.ModuleList()
        self.channel_splits = []
        for win, heads in window.items():
            self.conv_list.append(nn.Conv2d(heads * Ch, heads * Ch, win, padding=win // 2, groups=heads * Ch))
            self.channel_splits.append(heads * Ch)

    def forward(self, q, v, size):
        B, h, N, Ch = q.shape
        v_img = v.transpose(2, 3).reshape(B, h * Ch, *size)           # [B, (h Ch), H, W]
        parts = torch.split(v_img, self.channel_splits, dim=1)
        conv_v = torch.cat([conv(p) for conv, p in zip(self.conv_list, parts)], dim=1)
        return q * conv_v.reshape(B, h, Ch, N).transpose(2, 3)


class FactorAtt_ConvRelPosEnc(nn.Module):
    """Factorised attention: softmax over the TOKEN axis of k, (k^T v) first, then q (k^T v) (:333-393)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, shared_crpe=None):
        super().__init__()
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.crpe = shared_crpe

    def forward(self, x, size):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]                              # [B, h, N, Ch]
        ktv = k.softmax(dim=2).transpose(2, 3) @ v                    # [B, h, Ch, Ch]
        out = self.scale * (q @ ktv) + self.crpe(q, v, size=size)
        return self.proj(out.transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class MHCABlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=3, drop_path=0.0, qkv_bias=True, qk_scale=None, shared_cpe=None,
                 shared_crpe=None):
        super().__init__()
        self.cpe = shared_cpe      # registered here as well: the reference's state_dict repeats these keys per block
        self.crpe = shared_crpe
        self.factoratt_crpe = FactorAtt_ConvRelPosEnc(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                                      shared_crpe=shared_crpe)
        self.mlp = Mlp(dim, dim * mlp_ratio)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)

    def forward(self, x, size):
        if self.cpe is not None:
            x = self.cpe(x, size)
        x = x + self.drop_path(self.factoratt_crpe(self.norm1(x), size))
        return x + self.drop_path(self.mlp(self.norm2(x)))


class MHCAEncoder(nn.Module):
    def __init__(self, dim, num_layers=1, num_heads=8, mlp_ratio=3, drop_path_list=(), qk_scale=None,
                 crpe_window=None):
        super().__init__()
        self.num_layers = num_layers
        self.cpe = ConvPosEnc(dim, k=3)
        self.crpe = ConvRelPosEnc(Ch=dim // num_heads, h=num_heads, window=crpe_window or {3: 2, 5: 3, 7: 3})
        self.MHCA_layers = nn.ModuleList(
            MHCABlock(dim, num_heads=num_heads, mlp_ratio=mlp_ratio, drop_path=drop_path_list[idx], qk_scale=qk_scale,
                      shared_cpe=self.cpe, shared_crpe=self.crpe) for idx in range(num_layers))

    def forward(self, x, size):
        for layer in self.MHCA_layers:
            x = layer(x, size)
        return x.reshape(x.shape[0], *size, -1).permute(0, 3, 1, 2).contiguous()


class ResBlock(nn.Module):
    """1x1 (BN, Hardswish) -> depthwise 3x3 -> BN -> Hardswish -> 1x1 (BN) + identity (:482-532)."""

    def __init__(self, in_features, hidden_features=None, out_features=None):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.conv1 = Conv2d_BN(in_features, hidden_features, act_layer=nn.Hardswish)
        self.dwconv = nn.Conv2d(hidden_features, hidden_features, 3, 1, 1, bias=False, groups=hidden_features)
        self.norm = nn.BatchNorm2d(hidden_features)
        self.act = nn.Hardswish()
        self.conv2 = Conv2d_BN(hidden_features, out_features)
        nn.init.normal_(self.dwconv.weight, 0.0, math.sqrt(2.0 / 9.0))

    def forward(self, x):
        return x + self.conv2(self.act(self.norm(self.dwconv(self.conv1(x)))))


class MHCA_stage(nn.Module):
    def __init__(self, embed_dim, out_embed_dim, num_layers=1, num_heads=8, mlp_ratio=3, num_path=4,
                 drop_path_list=()):
        super().__init__()
        self.mhca_blks = nn.ModuleList(
            MHCAEncoder(embed_dim, num_layers, num_heads, mlp_ratio, drop_path_list=drop_path_list)
            for _ in range(num_path))
        self.InvRes = ResBlock(in_features=embed_dim, out_features=embed_dim)
        self.aggregate = Conv2d_BN(embed_dim * (num_path + 1), out_embed_dim, act_layer=nn.Hardswish)

    def forward(self, inputs):
        outs = [self.InvRes(inputs[0])]
        for x, encoder in zip(inputs, self.mhca_blks):
            outs.append(encoder(x.flatten(2).transpose(1, 2), size=tuple(x.shape[-2:])))
        return self.aggregate(torch.cat(outs, dim=1))


def dpr_generator(drop_path_rate, num_layers, num_stages):
    """Linear stochastic-depth schedule over all layers, cut per stage (:586-598)."""
    rates = [x.item() for x in torch.linspace(0, drop_path_rate, sum(num_layers))]
    out, cur = [], 0
    for i in range(num_stages):
        out.append(rates[cur:cur + num_layers[i]])
        cur += num_layers[i]
    return out


class MPViT(nn.Module):
    """forward(img [B,3,H,W]) -> list of 4 maps [B, C_{i+1}, H/2^{i+1}, W/2^{i+1}] (:601-730)."""

    def __init__(self, num_classes=80, in_chans=3, num_stages=4, num_layers=(1, 1, 1, 1), mlp_ratios=(8, 8, 4, 4),
                 num_path=(4, 4, 4, 4), embed_dims=(64, 128, 256, 512), num_heads=(8, 8, 8, 8), drop_path_rate=0.0,
                 norm_cfg=None, norm_eval=True, pretrained=None):
        super().__init__()
        self.num_classes, self.num_stages, self.norm_eval = num_classes, num_stages, norm_eval
        self.out_channels = [embed_dims[i + 1] if i + 1 < num_stages else embed_dims[i] for i in range(num_stages)]
        dpr = dpr_generator(drop_path_rate, list(num_layers), num_stages)
        self.stem = nn.Sequential(
            Conv2d_BN(in_chans, embed_dims[0] // 2, kernel_size=3, stride=1, pad=1, act_layer=nn.Hardswish),
            Conv2d_BN(embed_dims[0] // 2, embed_dims[0], kernel_size=3, stride=1, pad=1, act_layer=nn.Hardswish))
        self.patch_embed_stages = nn.ModuleList(
            Patch_Embed_stage(embed_dims[i], num_path=num_path[i], isPool=True) for i in range(num_stages))
        self.mhca_stages = nn.ModuleList(
            MHCA_stage(embed_dims[i], self.out_channels[i], num_layers[i], num_heads[i], mlp_ratios[i], num_path[i],
                       drop_path_list=dpr[i]) for i in range(num_stages))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        if isinstance(pretrained, str):
            self.load_pretrained(pretrained)

    def load_pretrained(self, path):
        ckpt = torch.load(path, map_location="cpu")
        self.load_state_dict(ckpt["model"] if "model" in ckpt else ckpt, strict=False)

    def forward_features(self, x):
        outs = []
        x = self.stem(x)
        for idx in range(self.num_stages):
            x = self.mhca_stages[idx](self.patch_embed_stages[idx](x))
            outs.append(x)
        return outs

    def forward(self, x):
        return self.forward_features(x)

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
        return self


_VARIANTS = {
    # name: (num_layers, embed_dims, mlp_ratio, drop_path_rate)     reference factories :743-870
    "mpvit_tiny": ((1, 2, 4, 1), (64, 96, 176, 216), 2, 0.0),
    "mpvit_xsmall": ((1, 2, 4, 1), (64, 128, 192, 256), 4, 0.0),
    "mpvit_small": ((1, 3, 6, 3), (64, 128, 216, 288), 4, 0.2),
    "mpvit_base": ((1, 3, 8, 3), (128, 224, 368, 480), 4, 0.4),
}


def _factory(name):
    layers, dims, ratio, dpr = _VARIANTS[name]

    def make(pretrained=None, **kwargs):
        """The reference factory loads `pretrained/<name>.pth` (a path on its authors' cluster for small/base) and
        fails without it; here the ImageNet checkpoint is optional: `pretrained=<path>`, or `$DD_MPVIT_PRETRAINED`,
        or `pretrained/<name>.pth` if it exists — a released DiffusionDepth checkpoint overrides it anyway."""
        model = MPViT(num_stages=4, num_path=(2, 3, 3, 3), num_layers=layers, embed_dims=dims,
                      mlp_ratios=(ratio,) * 4, num_heads=(8, 8, 8, 8), drop_path_rate=dpr, norm_eval=False, **kwargs)
        path = pretrained or os.environ.get("DD_MPVIT_PRETRAINED") or os.path.join("pretrained", name + ".pth")
        if os.path.isfile(path):
            model.load_pretrained(path)
        return model

    make.__name__ = name
    return make


mpvit_tiny = _factory("mpvit_tiny")
mpvit_xsmall = _factory("mpvit_xsmall")
mpvit_small = _factory("mpvit_small")
mpvit_base = _factory("mpvit_base")
