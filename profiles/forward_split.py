"""Where a native forward's GPU time goes (CUDA events, graph replay, BASELINE config 3): backbone / neck + FPN / encoder /
loop + decoder, through the engine's own entry points."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import dd_helpers
from oracle import restate
dev = torch.device('cuda:0')
T = int(os.environ.get("DD_STEPS", "20"))
m = dd_helpers.build_mirror('swinl', T).to(dev)
m.depth_head.check_range = False
m.depth_head.fp8_corrections = os.environ.get("AB_FP8", "1") == "1"
B, H, W = 4, 352, 1216
s = {k: v.to(dev) for k, v in restate.synthetic_sample(B, H, W).items()}
s["noise"] = restate.synthetic_noise(B, H, W).to(dev)
with torch.no_grad():
    for _ in range(2):
        m(s)
eng = next(iter(m.depth_head._engines.values()))
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
rgb, gt, noise = s["rgb"].contiguous(), s["gt"].contiguous(), s["noise"]
t_bb = timeit(lambda: eng.run_backbone(rgb))
feats = eng.run_backbone(rgb, want_feats=True)
t_cond_alone = timeit(lambda: eng.build_condition(feats))  # includes 4 NCHW -> plane transposes, runs the graph
def cond_only():
    eng.run_backbone(rgb); eng.build_condition(None)
t_bc = timeit(cond_only)
def upto_loop():
    eng.run_backbone(rgb); eng.build_condition(None); eng.encode(gt); eng.denoise_decode(None, noise)
t_all = timeit(upto_loop)
with torch.no_grad():
    t_fwd = timeit(lambda: m(s))
print(f"[{os.environ.get('AB_TAG','')}] backbone {t_bb:.2f} ms | neck+FPN {t_bc - t_bb:.2f} ms (alone, from NCHW feats: {t_cond_alone:.2f} ms) | encoder+loop+decoder {t_all - t_bc:.2f} ms "
      f"({(t_all - t_bc) / T:.3f} ms/step) | engine calls {t_all:.2f} ms | plugin forward {t_fwd:.2f} ms ({4e3 / t_fwd:.1f} maps/s)", flush=True)
