"""GN-apply + bilinear condition injection kernel: loop time and parity of the full loop vs the SIMT/debug path."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd
from diffusiondepth_b200.model.registry import HEADS
dev = torch.device('cuda:0')
torch.manual_seed(0)
head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64,128,256,512], inference_steps=20, num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
g = torch.Generator().manual_seed(0)
noise = torch.randn(4, 16, 176, 608, generator=g).to(dev); cond = torch.randn(4, 256, 88, 304, generator=g).abs().to(dev)
e = dd.DenoiseEngine("swin", 4, (176, 608), (88, 304), 20, dev, cuda_graph=True)
e.load_weights(head._engine_tensors()); e.set_schedule(*head.scheduler.fused_coefficients(20))
for _ in range(2): o = e.denoise_decode(cond, noise, want_logits=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): o = e.denoise_decode(cond, noise, want_logits=True)
e1.record(); torch.cuda.synchronize(); e.poll_status()
print(f"loop+decoder {e0.elapsed_time(e1)/5:.2f} ms ({e0.elapsed_time(e1)/100:.3f} ms/step)")
