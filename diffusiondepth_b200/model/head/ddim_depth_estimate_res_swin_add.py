"""`DDIMDepthEstimate_Swin_ADD` — the Swin-conditioned DDIM depth head WITHOUT the HAHI neck (reference
src/model/head/ddim_depth_estimate_res_swin_add.py:15-190): the four Swin stage outputs feed the FPN laterals
directly; the denoiser (bilinear-upsampled condition + convA/convB) is the one of the ADDHAHI head."""
from ..registry import HEADS
from ._ddim_head import DDIMHeadBase, _fpn_up


@HEADS.register_module()
class DDIMDepthEstimate_Swin_ADD(DDIMHeadBase):
    variant = "swin"
    has_neck = False
    fpn_in_channels = (192, 384, 768, 1536)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.convup_fp = _fpn_up()  # constructed by the reference (:49-59), never used in its forward
