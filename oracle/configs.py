"""TEST INFRASTRUCTURE ONLY — the BASELINE.json configurations and the golden cases derived from them."""
from argparse import Namespace

SEED_WEIGHTS = 7240   # reference src/config.py:42-45 default --seed
SEED_INPUTS = 0
SEED_NOISE = 1234

FAMILIES = {
    "res18": dict(backbone_module="mmbev_resnet", backbone_name="mmbev_res18", head_specify="DDIMDepthEstimate_Res"),
    "res50": dict(backbone_module="mmbev_resnet", backbone_name="mmbev_res50", head_specify="DDIMDepthEstimate_Res"),
    "swinl": dict(backbone_module="swin", backbone_name="swin_large_naive_nopretrain",
                  head_specify="DDIMDepthEstimate_Swin_ADDHAHI"),
    "swinl_add": dict(backbone_module="swin", backbone_name="swin_large_naive_nopretrain",
                      head_specify="DDIMDepthEstimate_Swin_ADD"),
    "mpvit_s": dict(backbone_module="mpvit", backbone_name="mpvit_small",
                    head_specify="DDIMDepthEstimate_MPVIT_ADDHAHI"),
    # the `*Vis` heads: same loop, `pred_inter` = the depth map decoded after EVERY step
    "res18_vis": dict(backbone_module="mmbev_resnet", backbone_name="mmbev_res18", head_specify="DDIMDepthEstimate_ResVis"),
    "swinl_vis": dict(backbone_module="swin", backbone_name="swin_large_naive_nopretrain",
                      head_specify="DDIMDepthEstimate_Swin_ADDHAHIVis"),
}

# name -> (family, T, batch, H, W).  C1..C5 = BASELINE.json configs[0..4]
CONFIGS = {
    "C1": ("res18", 5, 1, 228, 304),
    "C2": ("res50", 20, 8, 228, 304),
    "C3": ("swinl", 20, 4, 352, 1216),
    "C4": ("swinl", 20, 32, 352, 1216),
    "C5": ("swinl", 50, 64, 480, 640),
}

# golden cases (generated from the real reference by oracle/make_golden.py): name -> (family, T, batch, H, W)
GOLDEN = {
    "g_res18_c1": ("res18", 5, 1, 228, 304),        # BASELINE config 1 in full
    "g_res18_ragged": ("res18", 5, 2, 70, 106),     # odd latent (35 x 53): ragged tiles + adaptive pooling
    "g_swinl_small": ("swinl", 5, 1, 96, 160),      # Swin head on a small grid (CPU-cheap)
    "g_res50_c2": ("res50", 20, 1, 228, 304),       # image 0 of BASELINE config 2
    "g_swinl_c3": ("swinl", 20, 1, 352, 1216),      # image 0 of BASELINE config 3 / 4
    "g_swinl_c5": ("swinl", 50, 1, 480, 640),       # image 0 of BASELINE config 5 (50-step stress)
    "g_swinl_add_small": ("swinl_add", 5, 1, 96, 160),   # Swin head without the HAHI neck
    "g_mpvit_small": ("mpvit_s", 5, 1, 64, 112),    # MPViT-small + MPVIT_ADDHAHI head (cond at latent resolution)
}


# "trained-like" twins: same architectures, parameters perturbed by `trainedify` below, so that the regime released
# checkpoints live in (non-zero Swin relative-position tables, non-trivial BatchNorm running statistics, non-unit
# LayerNorm / GroupNorm affines) is pinned to the real reference too.  96x160 -> 24x40 tokens: padded (28x42) and
# shifted windows; 70x106 -> odd 35x53 latent, resampling FPN.
GOLDEN_TRAINED = {
    "g_swinl_small_trained": ("swinl", 5, 1, 96, 160),
    "g_res18_trained": ("res18", 5, 2, 70, 106),
    "g_mpvit_trained": ("mpvit_s", 5, 2, 70, 106),  # odd sizes on every MPViT level (35x53 .. 5x7), folded BN everywhere
    # Swin head on an odd image: patch embed pads 70x106 -> 72x108, condition map 18x27 under a 35x53 latent (align_corners
    # ratio exactly 0.5 but NOT the 2x layout: half of the outputs of the quad-based condition-injection kernel take their
    # per-tap path), every pyramid level odd (18x27, 9x14, 5x7, 3x4: resampling FPN, padded + shifted windows)
    "g_swinl_odd_trained": ("swinl", 5, 2, 70, 106),
    # `*Vis` heads (reference ..._res_vis.py / ..._swin_addHAHI_vis.py): the golden also holds `pred_inter` [T, B, 1, H, W]
    "g_res18_vis_trained": ("res18_vis", 5, 2, 70, 106),
    "g_swinl_vis_trained": ("swinl_vis", 5, 1, 96, 160),
}
SEED_TRAINED = 99


def trainedify(model, seed=SEED_TRAINED):
    """Deterministically move a freshly constructed model (reference or mirror: same module tree / key order) into a
    trained-like regime, in place.  CPU generator, fixed module order -> identical tensors wherever it runs."""
    import torch
    g = torch.Generator().manual_seed(seed)

    def rn(t, std, mean=0.0):
        t.copy_(torch.randn(t.shape, generator=g) * std + mean)

    def ru(t, lo, hi):
        t.copy_(torch.rand(t.shape, generator=g) * (hi - lo) + lo)

    with torch.no_grad():
        for name, mod in model.named_modules():
            cls = type(mod).__name__
            if cls == "BatchNorm2d":
                rn(mod.running_mean, 0.2); ru(mod.running_var, 0.5, 1.5); ru(mod.weight, 0.5, 1.5); rn(mod.bias, 0.2)
            elif cls in ("LayerNorm", "GroupNorm") and getattr(mod, "weight", None) is not None:
                ru(mod.weight, 0.7, 1.3); rn(mod.bias, 0.1)
            tab = getattr(mod, "relative_position_bias_table", None)
            if tab is not None:
                rn(tab, 0.5)
    return model


def make_args(family, steps):
    return Namespace(model_name="Diffusion_DCbase_", inference_steps=steps, num_train_timesteps=1000,
                     **FAMILIES[family])
