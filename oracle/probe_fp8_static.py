"""TEST INFRASTRUCTURE / design probe (CPU, no GPU): parity margin of the engine's FP8-CORRECTION scheme with the STATIC
power-of-two scales the kernels use (no per-tensor absmax pass), on the real BASELINE cases against the golden vectors of
the real reference.  Per 3x3 conv y = x * W (x: activations, W: weights), with xs = 16 x and ws = wscale * W:

    hi  = fp16(xs)            wh  = fp16(ws)                         -> hi  * wh   (fp16 MMA, K = 16)
    l8  = e4m3((xs - hi) * 512)      w8  = e4m3(ws / 512)            -> l8  * w8   (e4m3 MMA, K = 32, scales cancel)
    a8  = e4m3(xs / 4)               lw8 = e4m3((ws - wh) * 4)       -> a8  * lw8  (e4m3 MMA, K = 32, scales cancel)

all three into ONE fp32 accumulator.  `layers` selects which convs use it (the rest keep the exact 3-pass fp16 split).
Run:  python -m oracle.probe_fp8_static [C3|C5|small] [layers]     (layers: AB = convA+convB, ABN = + noise_embedding.3,
ALL = every conv but the raw-latent one)"""
import math, os, sys, time, torch, torch.nn.functional as F
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import dd_helpers
from oracle import restate, configs
torch.set_num_threads(int(os.environ.get("DD_THREADS", "8")))
CASES = {"C3": ("g_swinl_c3", 20), "C5": ("g_swinl_c5", 50), "small": ("g_swinl_small", 5)}
case = sys.argv[1] if len(sys.argv) > 1 else "C3"
sel = sys.argv[2] if len(sys.argv) > 2 else "AB"
FP8_LAYERS = {"AB": ("upsample_fuse.convA.conv", "upsample_fuse.convB.conv"),
              "ABN": ("upsample_fuse.convA.conv", "upsample_fuse.convB.conv", "noise_embedding.3"),
              "ALL": ("upsample_fuse.convA.conv", "upsample_fuse.convB.conv", "noise_embedding.3", "pred.0", "pred.3")}[sel]
gname, T = CASES[case]
g = dd_helpers.load_golden(gname)
m = dd_helpers.build_mirror('swinl', T)
sd = m.state_dict()
sample, noise = dd_helpers.inputs_for(g)
t = time.time()
with torch.no_grad():
    cond = restate.condition_features(sd, sample['rgb'], 'swin_large_naive_nopretrain')
print('cond', round(time.time() - t, 1), 's', tuple(cond.shape), flush=True)

def q16(t): return t.half().float()
def q8(t): return t.clamp(-448, 448).to(torch.float8_e4m3fn).float()
def wscale_of(w):
    a = w.abs().max().item()
    return 2.0 ** (math.floor(math.log2(32768.0 / a)) - 1) if a > 0 else 1.0   # engine.cu pack_layer

STAT = {}
def conv_emul(x, w, b, mode, name):
    sx = 16.0 if x.abs().max() < 3000 else 1.0
    sw = wscale_of(w)
    xs, ws = x * sx, w * sw
    xh = q16(xs); wh = q16(ws)
    if mode == 'fp8' and name in FP8_LAYERS:
        lo = xs - xh; wlo = ws - wh
        st = STAT.setdefault(name, [0.0, 0.0, 0.0])
        st[0] = max(st[0], xs.abs().max().item()); st[1] = max(st[1], (lo * 512).abs().max().item()); st[2] = max(st[2], (xs / 4).abs().max().item())
        y = F.conv2d(q8(lo * 512.0), q8(ws / 512.0), None, padding=1) + F.conv2d(q8(xs / 4.0), q8(wlo * 4.0), None, padding=1) \
            + F.conv2d(xh, wh, None, padding=1)
    else:
        xl = q16(xs - xh); wl = q16(ws - wh)
        y = F.conv2d(xl, wh, None, padding=1) + F.conv2d(xh, wl, None, padding=1) + F.conv2d(xh, wh, None, padding=1)
    return y / (sx * sw) + b.view(1, -1, 1, 1)

def denoiser(x_t, t, cond, mode):
    P = 'depth_head.model.'
    def cv(x, name): return conv_emul(x, sd[P + name + '.weight'], sd[P + name + '.bias'], mode, name)
    def gn(x, name): return torch.relu(F.group_norm(x, 4, sd[P + name + '.weight'], sd[P + name + '.bias'], 1e-5))
    feat = cond + sd[P + 'time_embedding.weight'][t][:, None, None]
    h = gn(cv(x_t, 'noise_embedding.0'), 'noise_embedding.1')
    ne = gn(cv(h, 'noise_embedding.3'), 'noise_embedding.4')
    up = F.interpolate(feat, size=ne.shape[-2:], mode='bilinear', align_corners=True)
    f = cv(cv(up + ne, 'upsample_fuse.convA.conv'), 'upsample_fuse.convB.conv')
    h = gn(cv(f, 'pred.0'), 'pred.1')
    return gn(cv(h, 'pred.3'), 'pred.4')

acp = restate.ddim_tables()
zg = torch.from_numpy(g['z']['logits'])
for mode in ('split3', 'fp8'):
    t0 = time.time()
    x = noise.clone()
    with torch.no_grad():
        for t in restate.ddim_timesteps(T):
            x = restate.ddim_step(denoiser(x, t, cond, mode), t, x, acp, T)
        z = restate.decode_logits(sd, x)
    dz = (dd_helpers.golden_view(g, 'logits', z) - zg).abs()
    print(f"{case} {sel} {mode}: {time.time() - t0:.0f} s  max|dz| vs reference golden {dz.max().item():.3e}  rms {dz.pow(2).mean().sqrt().item():.3e}", flush=True)
for k, v in STAT.items():
    print(f"  {k}: max|16x| {v[0]:.1f}  max|lo*512| {v[1]:.1f} (e4m3 max 448)  max|16x/4| {v[2]:.1f}")
