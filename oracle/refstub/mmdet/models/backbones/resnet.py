"""mmdet 2.x BasicBlock (Bottleneck kept as a constructor-compatible placeholder)."""
import torch.nn as nn
from mmcv.cnn import build_conv_layer, build_norm_layer
from mmcv.runner.base_module import BaseModule


class BasicBlock(BaseModule):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch',
                 with_cp=False, conv_cfg=None, norm_cfg=dict(type='BN'), dcn=None, plugins=None,
                 init_cfg=None):
        super().__init__(init_cfg)
        self.norm1_name, norm1 = build_norm_layer(norm_cfg, planes, postfix=1)
        self.norm2_name, norm2 = build_norm_layer(norm_cfg, planes, postfix=2)
        self.conv1 = build_conv_layer(conv_cfg, inplanes, planes, 3, stride=stride,
                                      padding=dilation, dilation=dilation, bias=False)
        self.add_module(self.norm1_name, norm1)
        self.conv2 = build_conv_layer(conv_cfg, planes, planes, 3, padding=1, bias=False)
        self.add_module(self.norm2_name, norm2)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = self.relu(getattr(self, self.norm1_name)(self.conv1(x)))
        out = getattr(self, self.norm2_name)(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        out = out + identity
        return self.relu(out)


class Bottleneck(BaseModule):
    expansion = 4

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError('Bottleneck blocks are not used by mmbev_res18/50/101')
