"""`DDIMDepthEstimate_Res` — ResNet-conditioned DDIM depth head (reference
src/model/head/ddim_depth_estimate_res.py:14-185): condition map at latent resolution, added to the noise
embedding (`feat = cond + temb + noise_embedding(x_t)`, :340), no upsample_fuse."""
import torch.nn as nn

from ..registry import HEADS
from ._ddim_head import FPN_DIM, DDIMHeadBase, _fpn_up


@HEADS.register_module()
class DDIMDepthEstimate_Res(DDIMHeadBase):
    variant = "res"
    fpn_in_channels = (64, 128, 256, 512)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.convup_fp = _fpn_up()  # constructed by the reference (:41-52), never used in its forward
