"""`DDIMDepthEstimate_Swin_ADDHAHI` — Swin-conditioned DDIM depth head with the HAHI neck (reference
src/model/head/ddim_depth_estimate_res_swin_addHAHI.py:15-185): condition map at half the latent
resolution, bilinear-upsampled (align_corners=True) and added to the noise embedding, then convA/convB."""
from ..necks.hahi import HAHIHeteroNeck
from ..registry import HEADS
from ._ddim_head import DDIMHeadBase


@HEADS.register_module()
class DDIMDepthEstimate_Swin_ADDHAHI(DDIMHeadBase):
    variant = "swin"
    has_neck = True
    fpn_in_channels = (192, 384, 768, 1536)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        c = list(self.fpn_in_channels)
        self.hahineck = HAHIHeteroNeck(in_channels=c, out_channels=c, embedding_dim=512,
                                       positional_encoding=dict(type='SinePositionalEncoding', num_feats=256),
                                       scales=[1, 1, 1, 1], cross_att=False, self_att=False, num_points=8)

    def _neck(self, fp):
        return self.hahineck(list(fp))
