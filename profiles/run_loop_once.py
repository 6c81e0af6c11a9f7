"""Profiling driver (run under ncu via gpurun): the hot path alone — Swin-variant denoise loop + decoder at the
BASELINE config-3 geometry (4 x [16,176,608] latents, cond [256,88,304], T=20) with random-init head weights.
No backbone, no CUDA graph, so every kernel of the path shows up as its own launch."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd  # noqa: E402
from diffusiondepth_b200.model.registry import HEADS  # noqa: E402

T = int(os.environ.get("DD_STEPS", "20"))
B = int(os.environ.get("DD_BATCH", "4"))
dev = torch.device("cuda:0")
torch.manual_seed(7240)
head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64, 128, 256, 512], inference_steps=T,
                        num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
g = torch.Generator().manual_seed(0)
noise = torch.randn(B, 16, 176, 608, generator=g).to(dev)
cond = torch.randn(B, 256, 88, 304, generator=g).abs().to(dev)
eng = dd.DenoiseEngine("swin", B, (176, 608), (88, 304), T, dev, cuda_graph=False)
eng.load_weights(head._engine_tensors())
eng.set_schedule(*head.scheduler.fused_coefficients(T))
for _ in range(int(os.environ.get("DD_REPEAT", "1"))):
    depth, _, _ = eng.denoise_decode(cond, noise)
torch.cuda.synchronize()
eng.poll_status()
print("launches", eng.last_launch_count, "depth mean", depth.mean().item())
