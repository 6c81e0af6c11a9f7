"""Depth <-> latent codec registry (reference src/model/ops/depth_transform.py).

Only the codec the shipped DDIM heads use is provided: `DeepDepthTransformWithUpsampling(hidden=16)`.
This module is the PARAMETER CONTAINER with the reference's key layout.  Inside a head's forward both directions run in
the CUDA engine: `t()` (encoder, :29-31; its value is only returned as `pred_init`) through `dd_encode`, `inv_t()`
(decoder, :33-35, on the hot path) through `dd_denoise_decode` / `dd_decode` (`DenoiseEngine.encode / .decode`).  The
torch expressions below exist for API parity when a caller invokes `t` / `inv_t` directly on a tensor (and for the
torch-op fallback of architectures the engine does not instantiate)."""
import torch
import torch.nn as nn

from ..registry import DEPTH_TRANSFORM


def _conv_block(cin, cout, k, stride, pad, bn=True, act=True):
    """conv(bias iff no BN) [+ BatchNorm2d] [+ LeakyReLU(0.2)] — key layout `0.weight`, `1.*`
    (reference src/model/common.py:45-60)."""
    mods = [nn.Conv2d(cin, cout, k, stride, pad, bias=not bn)]
    if bn:
        mods.append(nn.BatchNorm2d(cout))
    if act:
        mods.append(nn.LeakyReLU(0.2, inplace=True))
    return nn.Sequential(*mods)


@DEPTH_TRANSFORM.register_module()
class DeepDepthTransformWithUpsampling(nn.Module):
    def __init__(self, hidden=16, eps=1e-6):
        super().__init__()
        self.conv_transform = nn.Sequential(
            _conv_block(1, hidden, 3, 2, 1),
            _conv_block(hidden, hidden, 3, 1, 1, act=False),
            nn.Tanh())
        self.conv_inv_transform = nn.Sequential(
            nn.ConvTranspose2d(hidden, hidden, kernel_size=4, stride=2, padding=1),
            nn.BatchNorm2d(hidden),
            nn.ReLU(inplace=True),
            _conv_block(hidden, 1, 3, 1, 1, bn=False, act=False),
            nn.Sigmoid())
        self.eps = eps

    def t(self, depth):
        return self.conv_transform(depth)

    def inv_t(self, value):
        return 1.0 / self.conv_inv_transform(value).clamp(self.eps) - 1


@DEPTH_TRANSFORM.register_module()
class ReciprocalDepthTransform:
    """Parameter-free transform BaseDepthRefine builds by default before the DDIM heads replace it
    (reference mmbev_base_depth_refine.py:21, depth_transform.py:120-133)."""

    def __init__(self, linear=(1, 0), eps=1e-6):
        self.linear, self.eps = linear, eps

    def t(self, depth):
        return self.linear[0] / (1 + depth.clamp(0.)).clamp(self.eps) + self.linear[1]

    def inv_t(self, value):
        return self.linear[0] / (value - self.linear[1]).clamp(self.eps) - 1
