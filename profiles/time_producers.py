"""Where a forward's time goes on the GPU (CUDA events): backbone / neck / FPN / encoder (torch ops) vs engine."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import dd_helpers
from oracle import configs, restate
from diffusiondepth_b200.model._blocks import exact_fp32
dev = torch.device('cuda:0')
m = dd_helpers.build_mirror('swinl', 20).to(dev)
B, H, W = 4, 352, 1216
s = {k: v.to(dev) for k, v in restate.synthetic_sample(B, H, W).items()}
noise = restate.synthetic_noise(B, H, W).to(dev)
head = m.depth_head
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out
with torch.no_grad(), exact_fp32():
    t_bb, fp = timeit(lambda: m.depth_backbone(s['rgb']))
    t_neck, fp2 = timeit(lambda: head._neck(fp))
    t_fpn, cond = timeit(lambda: head._condition(fp2))
    t_enc, _ = timeit(lambda: head.depth_transform.t(s['gt']))
    cond = cond.contiguous()
    eng = head._engine(B, (176, 608), (88, 304), dev)
    t_eng, _ = timeit(lambda: eng.denoise_decode(cond, noise))
    t_all, _ = timeit(lambda: m({**s, 'noise': noise}))
print(f"backbone {t_bb:.1f} ms  neck {t_neck:.1f}  fpn {t_fpn:.1f}  encoder {t_enc:.2f}  engine(loop+dec) {t_eng:.1f}  full forward {t_all:.1f}")
