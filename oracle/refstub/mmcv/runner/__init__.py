from .base_module import BaseModule, ModuleList, Sequential  # noqa: F401


def force_fp32(*a, **k):
    def deco(f):
        return f
    return deco


auto_fp16 = force_fp32


def _load_checkpoint(filename, map_location=None, logger=None):
    import torch
    return torch.load(filename, map_location=map_location)


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None):
    ckpt = _load_checkpoint(filename, map_location)
    sd = ckpt.get('state_dict', ckpt.get('model', ckpt))
    model.load_state_dict(sd, strict=strict)
    return ckpt


def load_state_dict(module, state_dict, strict=False, logger=None):
    module.load_state_dict(state_dict, strict=strict)
