"""Small end-to-end run for compute-sanitizer: native Swin backbone + neck + FPN + 2-step loop + decoder at 64x96."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import dd_helpers
from oracle import restate
dev = torch.device("cuda:0")
m = dd_helpers.build_mirror("swinl", 2).to(dev)
m.depth_head.use_cuda_graph = False
s = {k: v.to(dev) for k, v in restate.synthetic_sample(1, 64, 96).items()}
s["noise"] = restate.synthetic_noise(1, 64, 96).to(dev)
with torch.no_grad():
    out = m(s)
torch.cuda.synchronize()
print("ok", out["pred"].shape, float(out["pred"].clamp(max=100).mean()))
r = dd_helpers.build_mirror("res18", 2).to(dev)
r.depth_head.use_cuda_graph = False
s = {k: v.to(dev) for k, v in restate.synthetic_sample(1, 36, 52).items()}
with torch.no_grad():
    out = r(s)
torch.cuda.synchronize()
print("ok res", out["pred"].shape)
