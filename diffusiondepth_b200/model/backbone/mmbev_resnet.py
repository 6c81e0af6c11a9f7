"""Stem-less BasicBlock ResNet that produces the four condition feature maps for the Res head
(reference src/model/backbone/mmbev_resnet.py:107-187; block = mmdet BasicBlock).  Runs once per image;
step-invariant, so it is outside the DDIM loop.  Keys: layers.{s}.{b}.{conv1,bn1,conv2,bn2,downsample}."""
import torch
import torch.nn as nn


class BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = downsample

    def forward(self, x):
        y = torch.relu_(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        skip = x if self.downsample is None else self.downsample(x)
        return torch.relu_(y + skip)


class ResNetForMMBEV(nn.Module):
    def __init__(self, numC_input, num_layer=(2, 2, 2), num_channels=None, stride=(2, 2, 2),
                 backbone_output_ids=None, norm_cfg=None, with_cp=False, block_type="Basic"):
        super().__init__()
        if block_type != "Basic":
            raise NotImplementedError("only the BasicBlock variants (mmbev_res18/50/101) are served")
        assert len(num_layer) == len(stride)
        if num_channels is None:
            num_channels = [numC_input * 2 ** (i + 1) for i in range(len(num_layer))]
        self.backbone_output_ids = range(len(num_layer)) if backbone_output_ids is None else backbone_output_ids
        stages, c = [], numC_input
        for n, width, s in zip(num_layer, num_channels, stride):
            # the skip of each stage's first block is a *biased* 3x3 strided conv with no norm (:128-130)
            blocks = [BasicBlock(c, width, stride=s, downsample=nn.Conv2d(c, width, 3, s, 1))]
            blocks += [BasicBlock(width, width) for _ in range(n - 1)]
            stages.append(nn.Sequential(*blocks))
            c = width
        self.layers = nn.Sequential(*stages)

    def forward(self, x):
        feats = []
        for i, stage in enumerate(self.layers):
            x = stage(x)
            if i in self.backbone_output_ids:
                feats.append(x)
        return feats


def _make(depths):
    return ResNetForMMBEV(3, num_layer=depths, num_channels=[64, 128, 256, 512], stride=[2, 2, 2, 2])


def mmbev_res18():
    return _make([2, 2, 2, 2])


def mmbev_res50():
    return _make([3, 4, 6, 3])


def mmbev_res101():
    return _make([3, 4, 23, 3])
