"""`DDIMDepthEstimate_ResVis` (reference src/model/head/ddim_depth_estimate_res_vis.py)."""
from ..registry import HEADS
from ._vis import VisMixin
from .ddim_depth_estimate_res import DDIMDepthEstimate_Res


@HEADS.register_module()
class DDIMDepthEstimate_ResVis(VisMixin, DDIMDepthEstimate_Res):
    pass
