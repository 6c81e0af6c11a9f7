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


def make_args(family, steps):
    return Namespace(model_name="Diffusion_DCbase_", inference_steps=steps, num_train_timesteps=1000,
                     **FAMILIES[family])
