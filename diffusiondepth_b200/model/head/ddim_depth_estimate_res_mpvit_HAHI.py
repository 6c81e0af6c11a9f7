"""`DDIMDepthEstimate_MPVIT_ADDHAHI` — the DDIM depth head on MPViT-small features (reference
src/model/head/ddim_depth_estimate_res_mpvit_HAHI.py:16-185): HAHI neck + FPN over channels 128/216/288/288 at
1/2 .. 1/16 of the image, the Swin heads' denoiser (upsample_fuse = bilinear(align_corners) + convA/convB; with the
condition map already at latent resolution the resize is the identity).

Neck, FPN, loop and decoder run on the engine like every other head: 216 and 288 are not multiples of the
tensor-core conv's 64-channel K chunk or of its N tiles, the partial chunks / tiles are completed with zeros by TMA's
out-of-bounds fill (convgen.cuh).  Only the MPViT backbone itself stays a torch module."""
from ..necks.hahi import HAHIHeteroNeck
from ..registry import HEADS
from ._ddim_head import DDIMHeadBase


@HEADS.register_module()
class DDIMDepthEstimate_MPVIT_ADDHAHI(DDIMHeadBase):
    variant = "swin"
    has_neck = True
    fpn_in_channels = (128, 216, 288, 288)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        c = list(self.fpn_in_channels)
        self.hahineck = HAHIHeteroNeck(in_channels=c, out_channels=c, embedding_dim=512,
                                       positional_encoding=dict(type='SinePositionalEncoding', num_feats=256),
                                       scales=[1, 1, 1, 1], cross_att=False, self_att=False, num_points=8)

    def _neck(self, fp):
        return self.hahineck(list(fp))
