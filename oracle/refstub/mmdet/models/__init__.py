from mmcv.utils import Registry
DETECTORS = Registry('detector')
BACKBONES = Registry('backbone')
NECKS = Registry('neck')
HEADS = Registry('head')
LOSSES = Registry('loss')
from . import backbones  # noqa: E402,F401
