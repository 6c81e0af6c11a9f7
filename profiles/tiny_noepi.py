"""16->64 classic kernel with the epilogue removed (DD_FP8_PROBE=2): mainloop-only tile rate."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd
from diffusiondepth_b200.model.registry import HEADS
dev = torch.device('cuda:0')
torch.manual_seed(0)
head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64,128,256,512], inference_steps=20, num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
os.environ["DD_SWAP_MASK"] = "0"; os.environ["DD_HALO_MASK"] = "6"
for probe in ("0", "2"):
    os.environ["DD_FP8_PROBE"] = probe
    e = dd.DenoiseEngine("swin", 4, (176, 608), (88, 304), 20, dev, cuda_graph=False)
    e.load_weights(head._engine_tensors()); e.set_schedule(*head.scheduler.fused_coefficients(20))
    line = f"probe={probe}"
    for cin, cout in [(16, 64), (64, 16), (256, 64)]:
        e.bench_conv(cin, cout, 5)
        line += f" | {cin}->{cout} {e.bench_conv(cin, cout, 40)*1e3:6.1f} us"
    print(line, flush=True)
