"""(The fp8 / ordering probe branches are compiled only with `DD_PROBES=1 python __graft_entry__.py --force`.)
SM cycles vs wall time per launch of the wide convs (DD_CLK_PROBE=1): single CTA vs CTA pair vs the fp8 probe."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DD_CLK_PROBE"] = "1"
import diffusiondepth_b200 as dd
from diffusiondepth_b200.model.registry import HEADS
dev = torch.device('cuda:0')
torch.manual_seed(0)
head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64,128,256,512], inference_steps=20, num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
for pair, fp8 in ((False, 0), (True, 0)):
    os.environ["DD_FP8_PROBE"] = str(fp8)
    e = dd.DenoiseEngine("swin", 4, (176, 608), (88, 304), 20, dev, cuda_graph=False, pair_wide=pair)
    e.load_weights(head._engine_tensors()); e.set_schedule(*head.scheduler.fused_coefficients(20))
    for cin, cout in [(64, 256), (256, 256)]:
        for iters in (1, 30):
            ms = e.bench_conv(cin, cout, iters)
            print(f"pair={pair} fp8={fp8} conv {cin}->{cout} iters={iters}: {ms*1e3:.1f} us", flush=True)
