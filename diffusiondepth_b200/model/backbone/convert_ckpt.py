"""Key / row-order conversion of the *official* Swin checkpoints (microsoft/Swin-Transformer layout) to the mmcv
layout this backbone uses (reference src/model/backbone/convert_ckpt.py:4-55).

Two things differ: names (`layers.*` -> `stages.*`, `attn.*` -> `attn.w_msa.*`, `mlp.fc1/fc2` ->
`ffn.layers.0.0 / ffn.layers.1`, `patch_embed.proj` -> `patch_embed.projection`) and the feature order of patch
merging: the official code concatenates the four 2x2 neighbours as [x0, x1, x2, x3] blocks (position-major, positions
ordered (0,0),(1,0),(0,1),(1,1)), this backbone unfolds channel-major with positions (0,0),(0,1),(1,0),(1,1) — so the
4C input axis of `downsample.reduction.weight` and `downsample.norm.{weight,bias}` is permuted accordingly."""
from collections import OrderedDict

import torch


def _to_unfold_order(t: torch.Tensor, axis: int) -> torch.Tensor:
    """[.., 4*C, ..] official (position-major, positions 0,1,2,3) -> unfold order (channel-major, positions 0,2,1,3)."""
    shape = list(t.shape)
    c = shape[axis] // 4
    v = t.movedim(axis, -1).reshape(*t.movedim(axis, -1).shape[:-1], 4, c)
    v = v[..., [0, 2, 1, 3], :].transpose(-1, -2).reshape(*v.shape[:-2], 4 * c)
    return v.movedim(-1, axis).contiguous()


_RENAMES = (("attn.", "attn.w_msa."), ("mlp.fc1.", "ffn.layers.0.0."), ("mlp.fc2.", "ffn.layers.1."), ("mlp.", "ffn."))


def swin_convert(ckpt):
    out = OrderedDict()
    for k, v in ckpt.items():
        if k.startswith("head"):
            continue
        nk, nv = k, v
        if k.startswith("layers"):
            if "attn." in k:
                nk = k.replace("attn.", "attn.w_msa.")
            elif "mlp." in k:
                for old, new in _RENAMES[1:]:
                    if old in k:
                        nk = k.replace(old, new)
                        break
            elif "downsample" in k:
                if "reduction." in k:
                    nv = _to_unfold_order(v, 1)
                elif "norm." in k:
                    nv = _to_unfold_order(v, 0)
            nk = nk.replace("layers", "stages", 1)
        elif k.startswith("patch_embed") and "proj" in k:
            nk = k.replace("proj", "projection")
        out[nk] = nv
    return out
