"""DDIM depth heads, registered under the reference's class names (reference src/model/head/__init__.py)."""
from .ddim_depth_estimate_res import DDIMDepthEstimate_Res  # noqa: F401
from .ddim_depth_estimate_res_swin_add import DDIMDepthEstimate_Swin_ADD  # noqa: F401
from .ddim_depth_estimate_res_swin_addHAHI import DDIMDepthEstimate_Swin_ADDHAHI  # noqa: F401
from .ddim_depth_estimate_res_vis import DDIMDepthEstimate_ResVis  # noqa: F401
from .ddim_depth_estimate_res_swin_addHAHI_vis import DDIMDepthEstimate_Swin_ADDHAHIVis  # noqa: F401
from .ddim_depth_estimate_res_mpvit_HAHI import DDIMDepthEstimate_MPVIT_ADDHAHI  # noqa: F401
