"""Mixin for the `*Vis` heads: additionally return every intermediate latent decoded to depth (`pred_inter`, a list of
T maps) — reference src/model/head/ddim_depth_estimate_res_swin_addHAHI_vis.py:130-149 (`inv_t` of every entry of the
pipeline's `image_list`, :289-304).  Same fully native path as the plain heads (backbone, neck, FPN, loop); the engine
decodes after every step inside the captured CUDA graph (`dd_denoise_decode_steps`), so a Vis forward is the plain
forward + T small decoder launches — no per-step host round trip, no NCHW<->NHWC traffic between steps."""


class VisMixin:
    return_intermediates = True
