"""In-situ kernel timeline of the hot path (CUPTI activity trace through torch.profiler — NOT ncu: kernels run back to back
inside the replayed CUDA graph, at the sustained clock, with warm L2).  Prints, per kernel name, launches / total time /
mean duration inside ONE loop + decoder replay, and the idle time between consecutive kernels (end -> next start), i.e.
what a serialised ncu launch list cannot show: how much of a graph-replayed DDIM step is spent between kernels.
DD_FULL=1: the whole plugin forward (native backbone + neck + FPN + loop) instead of the loop alone."""
import collections
import json
import os
import re
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd  # noqa: E402
from diffusiondepth_b200.model.registry import HEADS  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
T = int(os.environ.get("DD_STEPS", "20"))
full = os.environ.get("DD_FULL", "0") == "1"

if full:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import dd_helpers  # noqa: E402
    from oracle import restate  # noqa: E402  (input generation only)
    net = dd_helpers.build_mirror("swinl", T).to(dev)
    net.depth_head.check_range = False
    sample = {k: v.to(dev) for k, v in restate.synthetic_sample(4, 352, 1216).items()}
    sample["noise"] = restate.synthetic_noise(4, 352, 1216).to(dev)

    def run():
        with torch.no_grad():
            return net(sample)["pred"]
else:
    head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64, 128, 256, 512], inference_steps=T,
                            num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
    eng = dd.DenoiseEngine("swin", 4, (176, 608), (88, 304), T, dev, cuda_graph=True,
                           fp8_corr=os.environ.get("DD_EXACT", "0") != "1")
    eng.load_weights(head._engine_tensors())
    eng.set_schedule(*head.scheduler.fused_coefficients(T))
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(4, 16, 176, 608, generator=g).to(dev)
    cond = torch.randn(4, 256, 88, 304, generator=g).abs().to(dev)

    def run():
        return eng.denoise_decode(cond, noise)

for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        run()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel" and "dur" in e]
ev.sort(key=lambda e: e["ts"])
half = len(ev) // 2
ev = ev[half:]  # the second replay
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
busy = sum(e["dur"] for e in ev)
agg = collections.OrderedDict()
gaps = collections.OrderedDict()
for i, e in enumerate(ev):
    n = re.sub(r"\(.*", "", e["name"])
    n = re.sub(r"^void |dd::", "", n)[:70]
    a = agg.setdefault(n, [0, 0.0, []])
    a[0] += 1
    a[1] += e["dur"]
    a[2].append(e["dur"])
    if i + 1 < len(ev):
        gp = gaps.setdefault(n, [0, 0.0])
        gp[0] += 1
        gp[1] += max(0.0, ev[i + 1]["ts"] - (e["ts"] + e["dur"]))
out = {"kernels": len(ev), "span_us": span, "busy_us": busy, "idle_us": span - busy, "idle_frac": (span - busy) / span,
       "T": T, "full_forward": full, "per_kernel": []}
print(f"{len(ev)} kernels in one replay: span {span/1e3:.2f} ms, kernels {busy/1e3:.2f} ms, between kernels {(span-busy)/1e3:.2f} ms "
      f"({100*(span-busy)/span:.1f} %)")
for n, (c, t, durs) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gc, gt = gaps.get(n, [0, 0.0])
    sd = sorted(durs)
    med = sd[len(sd) // 2]
    # launches of one name alternate between call sites inside a step (convA / convB, the two 64-channel GN-apply passes)
    even = sorted(durs[0::2])
    odd = sorted(durs[1::2]) or [0.0]
    out["per_kernel"].append({"kernel": n, "launches": c, "total_us": t, "mean_us": t / c, "median_us": med, "min_us": sd[0],
                              "max_us": sd[-1], "median_even_us": even[len(even) // 2], "median_odd_us": odd[len(odd) // 2],
                              "gap_after_mean_us": gt / max(gc, 1)})
    print(f"{c:5d} x {t/c:8.1f} us = {t/1e3:8.2f} ms  med {med:7.1f} min {sd[0]:7.1f} max {sd[-1]:7.1f} even/odd {even[len(even)//2]:7.1f}/{odd[len(odd)//2]:7.1f}  {n}")
if os.environ.get("DD_DUMP"):
    # every launch of the replay in order: duration, grid, block, name (CUPTI's launch record)
    with open(os.environ["DD_DUMP"], "w") as f:
        for i, e in enumerate(ev):
            ar = e.get("args", {})
            f.write(f"{i:4d} {e['dur']:9.1f} us  grid {ar.get('grid')} block {ar.get('block')}  "
                    f"{re.sub(r'^void |dd::', '', re.sub(r'[(].*', '', e['name']))[:90]}\n")
dst = os.environ.get("DD_OUT")
if dst:
    json.dump(out, open(dst, "w"), indent=1)
