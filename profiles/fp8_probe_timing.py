"""(The fp8 / ordering probe branches are compiled only with `DD_PROBES=1 python __graft_entry__.py --force`.)
Timing-only probe: convA (256->256) with the two correction products issued as FP8 MMAs (DD_FP8_PROBE=1) vs the
production 3 x fp16 sequence.  Results of the probe mode are garbage by construction; only the kernel time matters."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd
from diffusiondepth_b200.model.registry import HEADS
dev = torch.device('cuda:0')
torch.manual_seed(0)
head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64,128,256,512], inference_steps=20, num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
eng = dd.DenoiseEngine("swin", 4, (176, 608), (88, 304), 20, dev, cuda_graph=False)
eng.load_weights(head._engine_tensors()); eng.set_schedule(*head.scheduler.fused_coefficients(20))
g = torch.Generator().manual_seed(0)
noise = torch.randn(4, 16, 176, 608, generator=g).to(dev); cond = torch.randn(4, 256, 88, 304, generator=g).abs().to(dev)
eng.denoise_decode(cond, noise); torch.cuda.synchronize()
P = 4 * 176 * 608
for rep in range(2):
    for probe in (False, True):
        if probe: os.environ["DD_FP8_PROBE"] = "1"
        else: os.environ.pop("DD_FP8_PROBE", None)
        for cin, cout in [(256, 256), (64, 256)]:
            ms = eng.bench_conv(cin, cout, 40)
            print(f"fp8_probe={probe} conv {cin}->{cout}: {ms*1e3:7.1f} us  ({2.0*P*cout*9*cin/(ms*1e-3)/1e12:6.1f} TF algorithmic)", flush=True)
