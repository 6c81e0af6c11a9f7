"""On-disk contract of the reference around the hot path: checkpoint ingestion and the KITTI
16-bit PNG result writer.  Host-side glue only."""
import os

import numpy as np
import torch


def load_reference_checkpoint(net: torch.nn.Module, path: str, map_location="cpu"):
    """`model_%05d.pt` as written by reference src/main.py:269-283 ({'net': state_dict, 'args': ...}); loaded the way
    `test()` does (:418-432): strict=False, unexpected keys reported, **missing keys are an error**."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"file not found: {path}")
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    state = ckpt["net"] if isinstance(ckpt, dict) and "net" in ckpt else ckpt
    state = {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}
    missing, unexpected = net.load_state_dict(state, strict=False)
    if unexpected:
        print("Unexpected keys :", unexpected)
    if missing:
        raise KeyError(f"Missing keys : {missing}")
    return ckpt.get("args") if isinstance(ckpt, dict) else None


def load_official_swin(backbone: torch.nn.Module, path: str, map_location="cpu"):
    """ImageNet Swin checkpoint in the official layout -> this backbone (reference swin.py:715-752)."""
    from .model.backbone.convert_ckpt import swin_convert
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    state = ckpt.get("state_dict", ckpt.get("model", ckpt))
    state = swin_convert(state)
    state = {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}
    own = backbone.state_dict()
    for k in [k for k in state if "relative_position_bias_table" in k and k in own]:
        src, dst = state[k], own[k]
        if src.shape != dst.shape and src.shape[1] == dst.shape[1]:  # window-size change: bicubic resize of the table
            s1, s2 = int(src.shape[0] ** 0.5), int(dst.shape[0] ** 0.5)
            t = torch.nn.functional.interpolate(src.permute(1, 0).reshape(1, -1, s1, s1), size=(s2, s2), mode="bicubic")
            state[k] = t.view(dst.shape[1], -1).permute(1, 0).contiguous()
    return backbone.load_state_dict(state, strict=False)


def depth_to_kitti_png(pred: torch.Tensor) -> np.ndarray:
    """`output['pred']` [B,1,H,W] -> the uint16 image reference summary/diffusion_dcbase_summary.py:166-186 writes
    (first image, clamp(min=0), depth * 256)."""
    d = torch.clamp(pred.detach(), min=0)[0, 0].cpu().numpy()
    return (d * 256.0).astype(np.uint16)


def save_kitti_png(pred: torch.Tensor, path: str):
    from PIL import Image
    Image.fromarray(depth_to_kitti_png(pred)).save(path)
