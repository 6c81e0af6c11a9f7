"""CTA-pair (tcgen05 cta_group::2) conv vs the single-CTA halo kernel: correctness on odd shapes, then timing on C3."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd
from diffusiondepth_b200.model.registry import HEADS
dev = torch.device('cuda:0')
eng = dd.DenoiseEngine('swin', 1, (8, 16), (4, 8), 2, dev, cuda_graph=False, pair_wide=True)
for (cin, cout) in [(64, 256), (256, 256)]:
    for (B, H, W) in [(2, 24, 40), (1, 16, 8), (1, 13, 21), (1, 16, 24), (2, 57, 76)]:
        g = torch.Generator().manual_seed(cin * 1000 + cout + H)
        x = torch.randn(B, cin, H, W, generator=g).to(dev) * 3
        w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).to(dev); b = torch.randn(cout, generator=g).to(dev)
        y = eng.conv3x3(x, w, b); torch.cuda.synchronize()
        ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
        err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
        print(f"pair conv {cin:3d}->{cout:3d} {B}x{H}x{W}: rel err {err:.2e}", flush=True)
torch.manual_seed(0)
head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64,128,256,512], inference_steps=20, num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
g = torch.Generator().manual_seed(0)
noise = torch.randn(4, 16, 176, 608, generator=g).to(dev); cond = torch.randn(4, 256, 88, 304, generator=g).abs().to(dev)
outs = {}
for pair in (False, True):
    e = dd.DenoiseEngine("swin", 4, (176, 608), (88, 304), 20, dev, cuda_graph=True, pair_wide=pair)
    e.load_weights(head._engine_tensors()); e.set_schedule(*head.scheduler.fused_coefficients(20))
    for _ in range(2): o = e.denoise_decode(cond, noise, want_logits=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): o = e.denoise_decode(cond, noise, want_logits=True)
    e1.record(); torch.cuda.synchronize()
    e.poll_status()
    outs[pair] = o[2]
    print(f"pair={pair}: loop+decoder {e0.elapsed_time(e1)/3:.2f} ms ({e0.elapsed_time(e1)/60:.3f} ms/step)")
    P = 4 * 176 * 608
    for cin, cout in [(64, 256), (256, 256)]:
        ms = e.bench_conv(cin, cout, 30)
        print(f"   conv {cin:3d}->{cout:3d}: {ms*1e3:7.1f} us  {2.0*P*cout*9*cin/(ms*1e-3)/1e12:6.1f} TF")
print("max |dz| pair vs single over the full 20-step loop:", (outs[True] - outs[False]).abs().max().item())
