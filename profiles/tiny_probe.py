"""Kernel choice for the two tiny layers (sid 0: 16->64, sid 4: 64->16): classic / halo / swap, timed alone (bench_conv)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd
from diffusiondepth_b200.model.registry import HEADS
dev = torch.device('cuda:0')
torch.manual_seed(0)
head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64,128,256,512], inference_steps=20, num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
for name, swap, halo in (("classic", 0b01000, 0b00110), ("halo", 0b01000, 0b10111), ("swap", 0b11001, 0b00110)):
    os.environ["DD_SWAP_MASK"] = str(swap); os.environ["DD_HALO_MASK"] = str(halo)
    e = dd.DenoiseEngine("swin", 4, (176, 608), (88, 304), 20, dev, cuda_graph=False)
    e.load_weights(head._engine_tensors()); e.set_schedule(*head.scheduler.fused_coefficients(20))
    line = f"{name:8s}"
    for cin, cout in [(16, 64), (64, 16), (256, 64)]:
        e.bench_conv(cin, cout, 5)
        line += f" | {cin}->{cout} {e.bench_conv(cin, cout, 40)*1e3:6.1f} us"
    print(line, flush=True)
