"""Small building blocks with the parameter names the reference's mmcv wrappers produce, so the mirror
loads reference checkpoints key-for-key, plus the fp32-exact execution context the
step-invariant producers (backbone / neck / FPN) run under."""
import contextlib

import torch
import torch.nn as nn


class ConvModule(nn.Module):
    """conv -> [BatchNorm2d as `.bn`] -> [ReLU]; bias iff no norm; kaiming-normal(fan_out) conv init —
    the semantics of mmcv.cnn.ConvModule the reference relies on (hahi.py:54-97, head :328-329)."""

    def __init__(self, cin, cout, k, padding=0, stride=1, norm=True, act=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=not norm)
        if norm:
            self.bn = nn.BatchNorm2d(cout)
        self.with_norm, self.with_act = norm, act
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode="fan_out", nonlinearity="relu")
        if self.conv.bias is not None:
            nn.init.zeros_(self.conv.bias)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.bn(x)
        return torch.relu_(x) if self.with_act else x


class DropPath(nn.Module):
    """Stochastic depth; identity in eval (the only mode the engine serves)."""

    def __init__(self, p=0.0):
        super().__init__()
        self.p = p

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        keep = 1.0 - self.p
        return x * x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep) / keep


@contextlib.contextmanager
def exact_fp32():
    """cuDNN/cuBLAS default to TF32 for fp32 convs/matmuls on GPU, which alone breaks the 1e-3 parity bar
    (a 4.9e-4 relative perturbation of the condition features gives 4.5e-3 on the output)."""
    c, m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = c, m
