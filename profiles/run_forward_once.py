"""Profiling driver: ONE full plugin forward (native Swin-L backbone + HAHI neck + FPN + T-step loop + decoder)
at BASELINE config 3 (4 x 352 x 1216), no CUDA graph so every kernel is its own launch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import dd_helpers
from oracle import restate
T = int(os.environ.get("DD_STEPS", "20")); B = int(os.environ.get("DD_BATCH", "4"))
dev = torch.device("cuda:0")
m = dd_helpers.build_mirror("swinl", T).to(dev)
m.depth_head.use_cuda_graph = False
s = {k: v.to(dev) for k, v in restate.synthetic_sample(B, 352, 1216).items()}
s["noise"] = restate.synthetic_noise(B, 352, 1216).to(dev)
with torch.no_grad():
    for _ in range(int(os.environ.get("DD_REPEAT", "1"))):
        out = m(s)
torch.cuda.synchronize()
eng = next(iter(m.depth_head._engines.values()))
print("launches", eng.last_launch_count, "pred mean", out["pred"].clamp(max=100).mean().item())
