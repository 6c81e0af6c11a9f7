from mmdet.models import HEADS, LOSSES, BACKBONES, NECKS, DETECTORS  # noqa: F401


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)
