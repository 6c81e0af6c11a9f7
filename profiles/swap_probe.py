"""Which narrow layers (sid 0: 16->64, 3: 256->64, 4: 64->16) should run on the swapped-operand kernel: correctness of
the quarter-local epilogue on odd shapes, then loop time for each mask (DD_SWAP_MASK bit = shape id)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd
from diffusiondepth_b200.model.registry import HEADS
dev = torch.device('cuda:0')
eng = dd.DenoiseEngine('swin', 1, (8, 16), (4, 8), 2, dev, cuda_graph=False, swap_narrow=True)
for (cin, cout) in [(16, 64), (256, 64), (64, 16)]:
    for (B, H, W) in [(2, 24, 40), (1, 16, 16), (1, 13, 21), (1, 5, 9), (2, 57, 76)]:
        g = torch.Generator().manual_seed(cin * 1000 + cout + H)
        x = torch.randn(B, cin, H, W, generator=g).to(dev) * 3
        w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).to(dev); b = torch.randn(cout, generator=g).to(dev)
        y = eng.conv3x3(x, w, b); torch.cuda.synchronize()
        ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
        err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
        print(f"swap conv {cin:3d}->{cout:3d} {B}x{H}x{W}: rel err {err:.2e}", flush=True)
torch.manual_seed(0)
head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64,128,256,512], inference_steps=20, num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
g = torch.Generator().manual_seed(0)
noise = torch.randn(4, 16, 176, 608, generator=g).to(dev); cond = torch.randn(4, 256, 88, 304, generator=g).abs().to(dev)
outs = {}
for mask in (0b01000, 0b01001, 0b11000, 0b11001, 0b00000):
    os.environ["DD_SWAP_MASK"] = str(mask)
    e = dd.DenoiseEngine("swin", 4, (176, 608), (88, 304), 20, dev, cuda_graph=True)
    e.load_weights(head._engine_tensors()); e.set_schedule(*head.scheduler.fused_coefficients(20))
    for _ in range(2): o = e.denoise_decode(cond, noise, want_logits=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): o = e.denoise_decode(cond, noise, want_logits=True)
    e1.record(); torch.cuda.synchronize()
    e.poll_status()
    outs[mask] = o[2]
    line = f"mask={mask:05b}: loop+decoder {e0.elapsed_time(e1)/3:.2f} ms ({e0.elapsed_time(e1)/60:.3f} ms/step)"
    for cin, cout in [(16, 64), (256, 64), (64, 16)]:
        line += f" | {cin}->{cout} {e.bench_conv(cin, cout, 30)*1e3:6.1f} us"
    print(line, flush=True)
for m in outs: print(f"max |dz| mask {m:05b} vs 01000:", (outs[m] - outs[0b01000]).abs().max().item())
