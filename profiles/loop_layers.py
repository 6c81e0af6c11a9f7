import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd
from diffusiondepth_b200.model.registry import HEADS
dev = torch.device('cuda:0')
torch.manual_seed(0)
head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64,128,256,512], inference_steps=20, num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
eng = dd.DenoiseEngine("swin", 4, (176, 608), (88, 304), 20, dev, cuda_graph=True)
eng.load_weights(head._engine_tensors()); eng.set_schedule(*head.scheduler.fused_coefficients(20))
g = torch.Generator().manual_seed(0)
noise = torch.randn(4, 16, 176, 608, generator=g).to(dev); cond = torch.randn(4, 256, 88, 304, generator=g).abs().to(dev)
for _ in range(2): eng.denoise_decode(cond, noise)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): eng.denoise_decode(cond, noise)
e1.record(); torch.cuda.synchronize()
print(f"loop+decoder (graph): {e0.elapsed_time(e1)/3:.2f} ms  -> {e0.elapsed_time(e1)/3/20:.3f} ms/step")
P = 4 * 176 * 608
tot = 0
for cin, cout, n in [(16, 64, 1), (64, 256, 1), (256, 256, 2), (256, 64, 1), (64, 16, 1)]:
    ms = eng.bench_conv(cin, cout, 30)
    tf = 2.0 * P * cout * 9 * cin / (ms * 1e-3) / 1e12
    tot += ms * n
    print(f"conv {cin:3d}->{cout:3d}: {ms*1e3:7.1f} us/launch  {tf:6.1f} TF algorithmic  (x{n}/step)")
print(f"convs per step: {tot:.3f} ms")
