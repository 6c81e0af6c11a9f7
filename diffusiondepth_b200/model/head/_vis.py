"""Mixin for the `*Vis` heads: additionally return every intermediate latent decoded to depth
(`pred_inter`, a list of T maps) — reference ..._swin_addHAHI_vis.py:130-149,289-304.  The fused CUDA loop
keeps intermediates on chip-side buffers only, so this path drives the engine one step at a time
(denoiser operator + collapsed DDIM update + decoder per step)."""
import torch

from .._blocks import exact_fp32


class VisMixin:
    def can_run_backbone(self, backbone, img) -> bool:
        return False  # the step-wise Vis path takes backbone features from the torch modules

    def forward(self, fp, depth_map, depth_mask, gt_depth_map=None, return_loss=False, noise=None, **kwargs):
        with torch.no_grad(), exact_fp32():
            gt_map_t = self.depth_transform.t(gt_depth_map)
            cond = self._condition(self._neck(fp)).contiguous()
        B = cond.shape[0]
        x = self._draw_noise((B, *gt_map_t.shape[-3:]), cond.device, cond.dtype, noise)
        eng = self._engine(B, tuple(gt_map_t.shape[-2:]), tuple(cond.shape[-2:]), cond.device)
        ts, cx, ce = self.scheduler.fused_coefficients(self.diffusion_inference_steps)
        inter = []
        for t, a, b in zip(ts, cx, ce):
            eps = eng.denoiser_forward(cond, x, t)
            x = (a * x + b * eps).contiguous()
            inter.append(eng.decode(x)[0])
        if self.check_range:
            eng.poll_status()
        self.last_latent = x
        ddim_loss = self._ddim_loss(cond, x) if (self.eval_ddim_loss or self.training) else x.new_zeros(())
        return {'pred': inter[-1], 'pred_init': gt_map_t, 'blur_depth_t': gt_map_t, 'ddim_loss': ddim_loss,
                'gt_map_t': gt_map_t, 'pred_uncertainty': None, 'pred_inter': inter, 'weight_map': None,
                'guidance': None, 'offset': None, 'aff': None, 'gamma': None, 'confidence': None}
