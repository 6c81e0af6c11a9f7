"""TEST INFRASTRUCTURE ONLY — import the *unmodified* reference (`/root/reference/src/model`)
under the in-repo stubs of mmcv/mmdet/mmdet3d/apex (oracle/refstub).

Only usable where `/root/reference` exists (the build container); the GPU box never has it.
Used by `oracle/make_golden.py` (fixture generation) and by the `not gpu` tests that pin
`oracle/restate.py` against the real reference.  Nothing under `diffusiondepth_b200/` imports this.
"""
import os
import sys
import types
from argparse import Namespace

REF_SRC = os.environ.get("DD_REFERENCE_SRC", "/root/reference/src")
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstub")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "model"))


def _activate():
    if not available():
        raise RuntimeError(f"reference sources not found under {REF_SRC}")
    for p in (_STUB, REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the product mirror is also importable as top-level `model`; make sure the
    # reference wins inside this process.
    m = sys.modules.get("model")
    if m is not None and not getattr(m, "__file__", "").startswith(REF_SRC):
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]


def make_args(backbone_module, backbone_name, head_specify, inference_steps=20,
              num_train_timesteps=1000):
    """The subset of reference src/config.py flags the model reads (config.py:116-141)."""
    return Namespace(backbone_module=backbone_module, backbone_name=backbone_name,
                     head_specify=head_specify, inference_steps=inference_steps,
                     num_train_timesteps=num_train_timesteps, model_name='Diffusion_DCbase_')


def build_reference_model(args):
    """reference src/main.py:414 — `get_model(args)(args)`; returns the nn.Module in eval().

    The reference's MPViT factories `torch.load` an ImageNet checkpoint from a hard-coded path
    (backbone/mpvit.py:756,788,826,859) that exists nowhere but on its authors' cluster; while such a model is being
    constructed, loads of a missing `*mpvit_*.pth` return an empty {'model': {}} (load_state_dict(strict=False) of
    nothing), i.e. the network keeps its constructor initialisation — the weights are overwritten by the caller's
    load_state_dict(strict=True) anyway."""
    _activate()
    import importlib
    import torch
    model_pkg = importlib.import_module("model")
    cls = model_pkg.get(args)
    real_load = torch.load

    def load(f, *a, **k):
        if isinstance(f, str) and "mpvit_" in os.path.basename(f) and not os.path.isfile(f):
            return {"model": {}}
        return real_load(f, *a, **k)

    torch.load = load
    try:
        net = cls(args)
    finally:
        torch.load = real_load
    net.eval()
    return net


def reference_modules():
    """Handles to the reference's own sub-plugins (scheduler, codec registry, heads)."""
    _activate()
    import importlib
    ns = types.SimpleNamespace()
    ns.scheduling_ddim = importlib.import_module("model.diffusers.schedulers.scheduling_ddim")
    ns.depth_transform = importlib.import_module("model.ops.depth_transform")
    ns.head_swin = importlib.import_module("model.head.ddim_depth_estimate_res_swin_addHAHI")
    ns.head_res = importlib.import_module("model.head.ddim_depth_estimate_res")
    ns.swin = importlib.import_module("model.backbone.swin")
    ns.resnet = importlib.import_module("model.backbone.mmbev_resnet")
    ns.hahi = importlib.import_module("model.necks.hahi")
    ns.mpvit = importlib.import_module("model.backbone.mpvit")
    return ns
