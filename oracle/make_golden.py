"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the REAL reference (/root/reference, imported
unmodified under oracle/refstub).  Run in the build container:  python -m oracle.make_golden [case ...]

Weights: the product mirror's default construction under torch.manual_seed(7240) (same init distributions as
the reference's modules; its state_dict matches the reference key-for-key), loaded into the reference with
load_state_dict(strict=True).  Inputs/noise: oracle.restate.synthetic_sample / synthetic_noise.
Stored per case: decoder logits z, depth, final latent, condition map (sub-sampled where large), summary
statistics and a weight checksum so a consumer can tell whether it regenerated the same weights."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import configs, ref_import, reference_runner, restate  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def build_mirror(family, steps, trained=False):
    from diffusiondepth_b200.model import get
    args = configs.make_args(family, steps)
    torch.manual_seed(configs.SEED_WEIGHTS)
    m = get(args)(args).eval()
    return configs.trainedify(m) if trained else m


def weight_checksum(sd):
    """sum |w| in fp64 over the hot-path parameters + a few producer tensors."""
    keys = sorted(k for k in sd if k.startswith("depth_head.model.") or "conv_inv_transform" in k
                  or k.startswith("depth_head.conv_lateral") or k.endswith("relative_position_bias_table"))
    return float(sum(sd[k].double().abs().sum() for k in keys if sd[k].is_floating_point()))


def subsample(name, t, H, W):
    t = t.detach().float()
    if name in ("logits", "pred"):
        s = 1 if H * W <= 80000 else 4
        return t[..., ::s, ::s].contiguous(), s
    if name == "latent":
        s = 1 if H * W <= 20000 else (2 if H * W <= 80000 else 4)
        return t[..., ::s, ::s].contiguous(), s
    if name == "cond":
        return t[:, ::32, ::4, ::4].contiguous(), 4
    raise KeyError(name)


def generate(case):
    trained = case in configs.GOLDEN_TRAINED
    family, T, B, H, W = (configs.GOLDEN_TRAINED if trained else configs.GOLDEN)[case]
    t0 = time.time()
    mirror = build_mirror(family, T, trained)
    sd = {k: v.detach().clone() for k, v in mirror.state_dict().items()}
    ref = ref_import.build_reference_model(ref_import.make_args(
        configs.FAMILIES[family]["backbone_module"], configs.FAMILIES[family]["backbone_name"],
        configs.FAMILIES[family]["head_specify"], T))
    ref.load_state_dict(sd, strict=True)
    sample = restate.synthetic_sample(B, H, W, configs.SEED_INPUTS)
    noise = restate.synthetic_noise(B, H, W, configs.SEED_NOISE)
    torch.manual_seed(0)
    r = reference_runner.run_reference(ref, sample, noise)
    t_ref = time.time() - t0
    # pin the restatement against the reference on this very case
    o = restate.forward(sd, sample, configs.FAMILIES[family]["backbone_name"], T, noise)
    pm = restate.parity_metrics(o["logits"], r["logits"], o["pred"], r["pred"])
    lat_err = (o["latent"] - r["latent"]).abs().max().item() / max(r["latent"].abs().max().item(), 1e-30)
    cond_err = (o["cond"] - r["cond"]).abs().max().item() / max(r["cond"].abs().max().item(), 1e-30)
    print(f"[{case}] reference {t_ref:.1f}s; oracle-vs-reference: max|dz|={pm['max_dz']:.2e} rms={pm['rms_dz']:.2e} "
          f"latent rel {lat_err:.2e} cond rel {cond_err:.2e}", flush=True)
    arrays = {}
    for name in ("logits", "pred", "latent", "cond"):
        arr, stride = subsample(name, r[name], H, W)
        arrays[name] = arr.numpy()
        arrays[name + "_stride"] = np.int32(stride)
    if "pred_inter" in r:
        arrays["pred_inter"] = r["pred_inter"].float().numpy()
    z = r["logits"].double()
    arrays.update(
        meta=np.array([T, B, H, W], dtype=np.int32), family=np.array(family),
        weight_checksum=np.float64(weight_checksum(sd)),
        logits_mean=np.float64(z.mean()), logits_std=np.float64(z.std()), logits_absmax=np.float64(z.abs().max()),
        latent_std=np.float64(r["latent"].double().std()), latent_absmax=np.float64(r["latent"].abs().max()),
        cond_absmax=np.float64(r["cond"].abs().max()),
        frac_clamped=np.float64((r["pred"] >= 999998.0).double().mean()),
        oracle_max_dz=np.float64(pm["max_dz"]), output_keys=np.array(r["keys"]))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, case + ".npz"), **arrays)
    print(f"[{case}] wrote {case}.npz ({os.path.getsize(os.path.join(OUT, case + '.npz')) / 1e3:.0f} kB)", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 8)
    for c in (sys.argv[1:] or list(configs.GOLDEN) + list(configs.GOLDEN_TRAINED)):
        generate(c)
