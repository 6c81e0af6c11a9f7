"""TEST INFRASTRUCTURE ONLY — run the *real* reference model (oracle/ref_import.py) with an injected
initial latent and capture the internal tensors parity is judged on (condition map, final latent,
decoder logit).  Needs /root/reference; used by oracle/make_golden.py and tests/test_oracle_golden.py."""
import contextlib

import torch


@contextlib.contextmanager
def _inject_first_randn(noise):
    """The loop's initial latent is the first `torch.randn` draw of an eval forward
    (head :283; ddim_loss draws only afterwards, :209) — hand it our tensor, leave later draws alone."""
    real = torch.randn
    state = {"used": False}

    def fake(*size, **kw):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        if not state["used"] and shape == tuple(noise.shape):
            state["used"] = True
            return noise.clone().to(kw.get("device") or "cpu", kw.get("dtype") or noise.dtype)
        return real(*size, **kw)

    torch.randn = fake
    try:
        yield state
    finally:
        torch.randn = real


def run_reference(net, sample, noise):
    """-> dict(pred, logits, latent, cond, pred_init).  `net` = reference Diffusion_DCbase_Model in eval()."""
    cap = {}
    head = net.depth_head
    def grab_cond(m, a):
        cap.setdefault("cond", a[2].detach().clone())  # first denoiser call = first loop step

    def grab_latent(m, a):
        cap["latent"] = a[0].detach().clone()

    def grab_logits(m, a, o):
        cap["logits"] = o.detach().clone()

    hooks = [head.model.register_forward_pre_hook(grab_cond),
             head.depth_transform.conv_inv_transform.register_forward_pre_hook(grab_latent),
             head.depth_transform.conv_inv_transform[3].register_forward_hook(grab_logits)]
    try:
        with torch.no_grad(), _inject_first_randn(noise) as st:
            out = net(sample)
        assert st["used"], "the reference did not draw the initial latent with the expected shape"
    finally:
        for h in hooks:
            h.remove()
    r = dict(pred=out["pred"], logits=cap["logits"], latent=cap["latent"], cond=cap["cond"],
             pred_init=out["pred_init"], keys=sorted(out.keys()))
    if out.get("pred_inter") is not None:  # the `*Vis` heads: one decoded map per DDIM step
        r["pred_inter"] = torch.stack([p.detach() for p in out["pred_inter"]])
    return r
