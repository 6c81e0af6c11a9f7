"""A/B timing of the loop convolutions across library builds and probe modes (CUDA events + in-kernel clock probe; never
under ncu).  One subprocess per configuration because the library path / probe variables are read once per process.

  python profiles/ab_probe.py            # all configurations below
Configurations: the round-1 library (profiles/_ab/libddengine_r1.so, lone-lane TMA producers), the current product
library, and the -DDD_PROBES build with DD_PAIR_MASK / DD_FP8_PROBE variants (fp8 results are garbage: timing only)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffusiondepth_b200", "libddengine.so")
PROBES = LIB.replace(".so", "_probes.so")
R1 = os.path.join(ROOT, "profiles", "_ab", "libddengine_r1.so")

CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
import diffusiondepth_b200 as dd
from diffusiondepth_b200.model.registry import HEADS
dev = torch.device("cuda:0")
torch.manual_seed(0)
head = HEADS.build(dict(type="DDIMDepthEstimate_Swin_ADDHAHI", in_channels=[64,128,256,512], inference_steps=20,
                        num_train_timesteps=1000, depth_feature_dim=16, loss_cfgs=[], init_cfg=None)).eval().to(dev)
g = torch.Generator().manual_seed(0)
noise = torch.randn(4, 16, 176, 608, generator=g).to(dev); cond = torch.randn(4, 256, 88, 304, generator=g).abs().to(dev)
kw = dict(fp8_corr=os.environ.get("AB_FP8", "1") == "1") if "fp8_corr" in dd.DenoiseEngine.__init__.__code__.co_varnames else {}
e = dd.DenoiseEngine("swin", 4, (176, 608), (88, 304), 20, dev, cuda_graph=True, **kw)
e.load_weights(head._engine_tensors()); e.set_schedule(*head.scheduler.fused_coefficients(20))
tag = os.environ["AB_TAG"]
if os.environ.get("AB_LOOP") == "1":
    for _ in range(2): e.denoise_decode(cond, noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): e.denoise_decode(cond, noise)
    e1.record(); torch.cuda.synchronize()
    print(f"[{tag}] loop+decoder {e0.elapsed_time(e1)/3:.2f} ms = {e0.elapsed_time(e1)/60:.3f} ms/step", flush=True)
else:
    e.denoise_decode(cond, noise); torch.cuda.synchronize()
P = 4 * 176 * 608
shapes = [(256, 256), (64, 256)] + ([(256, 64), (64, 16), (16, 64)] if os.environ.get("AB_ALL") == "1" else [])
for cin, cout in shapes:
    ms = e.bench_conv(cin, cout, 30)
    print(f"[{tag}] conv {cin:3d}->{cout:3d}: {ms*1e3:7.1f} us  {2.0*P*cout*9*cin/(ms*1e-3)/1e12:6.1f} TF algorithmic", flush=True)
''' % ROOT

CONFIGS = [
    ("r1 lib (lone-lane TMA producers)", R1, dict(AB_LOOP="1", AB_ALL="1")),
    ("r2 lib (product build, default flags)", LIB, dict(AB_LOOP="1", AB_ALL="1")),
    ("r2 lib, exact 3-pass split (fp8_corr off)", LIB, dict(AB_LOOP="1", AB_ALL="1", AB_FP8="0")),
    ("probes: 256->256 single, 64->256 pair", PROBES, dict(DD_CLK_PROBE="1", DD_PAIR_MASK="2")),
    ("probes: pairs for both", PROBES, dict(DD_CLK_PROBE="1", DD_PAIR_MASK="6")),
    ("probes: single both, fp8 only (3 x K32 e4m3)", PROBES, dict(DD_CLK_PROBE="1", DD_PAIR_MASK="0", DD_FP8_PROBE="3")),
    ("probes: pair both, fp8 only (3 x K32 e4m3)", PROBES, dict(DD_CLK_PROBE="1", DD_PAIR_MASK="6", DD_FP8_PROBE="3")),
    ("probes: single both, 2 fp16 + 2 fp8", PROBES, dict(DD_CLK_PROBE="1", DD_PAIR_MASK="0", DD_FP8_PROBE="1")),
    ("probes: pair both, 2 fp16 + 2 fp8", PROBES, dict(DD_CLK_PROBE="1", DD_PAIR_MASK="6", DD_FP8_PROBE="1")),
]
ONLY = os.environ.get("AB_ONLY")  # comma-separated indices into CONFIGS
if ONLY:
    CONFIGS = [CONFIGS[int(i)] for i in ONLY.split(",")]
for tag, lib, env in CONFIGS:
    if not os.path.exists(lib):
        print(f"[{tag}] skipped: {lib} missing", flush=True)
        continue
    full = dict(os.environ, DD_ENGINE_LIB=lib, AB_TAG=tag, **env)
    r = subprocess.run([sys.executable, "-c", CHILD], env=full, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout)
    for line in r.stderr.splitlines():
        if "clk_probe" in line or "Error" in line or "error" in line:
            print(f"[{tag}] {line}")
    sys.stdout.flush()
