"""`DDIMDepthEstimate_Swin_ADDHAHIVis` (reference src/model/head/ddim_depth_estimate_res_swin_addHAHI_vis.py)."""
from ..registry import HEADS
from ._vis import VisMixin
from .ddim_depth_estimate_res_swin_addHAHI import DDIMDepthEstimate_Swin_ADDHAHI


@HEADS.register_module()
class DDIMDepthEstimate_Swin_ADDHAHIVis(VisMixin, DDIMDepthEstimate_Swin_ADDHAHI):
    pass
