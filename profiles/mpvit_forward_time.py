"""Times the MPViT-small plugin forward (DDIMDepthEstimate_MPVIT_ADDHAHI, T=20) at KITTI size on one GPU: backbone on the
engine vs the torch-module backbone (neck + FPN + loop + decoder on the engine in both)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dd_helpers as helpers  # noqa: E402
from oracle import configs, restate  # noqa: E402

B, H, W, T = int(os.environ.get("DD_B", 4)), 352, 1216, 20
dev = torch.device("cuda:0")
m = helpers.build_mirror("mpvit_s", T).to(dev)
sample = restate.synthetic_sample(B, H, W, configs.SEED_INPUTS)
sample["noise"] = restate.synthetic_noise(B, H, W, configs.SEED_NOISE)
sample = {k: v.to(dev) for k, v in sample.items()}
out = {}
REPS = int(os.environ.get("DD_REPS", 5))
for native in ((True,) if os.environ.get("DD_NATIVE_ONLY") else (True, False)):
    m.depth_head.native_backbone = native
    with torch.no_grad():
        for _ in range(min(3, REPS)):
            m(sample)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(REPS):
            o = m(sample)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    eng = next(reversed(m.depth_head._engines.values()))
    out["engine_backbone" if native else "torch_backbone"] = {
        "ms_per_forward": ms, "maps_per_s": B / ms * 1e3, "engine_launches": eng.last_launch_count,
        "pred_mean": float(o["pred"].clamp(max=1e3).mean())}
    if native:
        # backbone alone
        eng.run_backbone(sample["rgb"].contiguous().float())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(REPS):
            eng.run_backbone(sample["rgb"].contiguous().float())
        e1.record()
        torch.cuda.synchronize()
        out["engine_backbone"]["backbone_ms"] = e0.elapsed_time(e1) / REPS
print(json.dumps({"workload": f"MPViT-small, T={T}, {B}x{H}x{W}", **out}))
