"""TEST INFRASTRUCTURE / design probe (CPU, no GPU): would Winograd F(2x2, 3x3) on the two 256->256 convs keep parity?

F(2x2,3x3): a 4x4 input tile d and a 3x3 filter g give a 2x2 output tile  Y = A^T [ (G g G^T) (.) (B^T d B) ] A :
16 multiplies per 4 outputs instead of 36 (2.25x fewer MACs).  On the engine each of the 16 transform-domain positions
would be a GEMM over channels with the 3-pass fp16 split of its operands; the transforms are fp32 adds.  This emulation
does exactly that (input transform in fp32, transformed operands split hi/lo with the engine's scales, the three partial
products per position in fp32, output transform in fp32) for upsample_fuse.convA / convB on the real BASELINE case
against the golden of the real reference; every other conv keeps the exact 3-pass direct form.
Run:  python -m oracle.probe_winograd [C3|small]"""
import math, os, sys, time, torch, torch.nn.functional as F
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import dd_helpers
from oracle import restate
torch.set_num_threads(int(os.environ.get("DD_THREADS", "8")))
CASES = {"C3": ("g_swinl_c3", 20), "small": ("g_swinl_small", 5)}
case = sys.argv[1] if len(sys.argv) > 1 else "C3"
gname, T = CASES[case]
g = dd_helpers.load_golden(gname)
m = dd_helpers.build_mirror('swinl', T)
sd = m.state_dict()
sample, noise = dd_helpers.inputs_for(g)
with torch.no_grad():
    cond = restate.condition_features(sd, sample['rgb'], 'swin_large_naive_nopretrain')
BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float32)
G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float64)
AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float32)
def q16(t): return t.half().float()
def pow2(a, target): return 2.0 ** math.floor(math.log2(target / a)) if a > 0 else 1.0
WINO = ("upsample_fuse.convA.conv", "upsample_fuse.convB.conv")
STAT = {}
def conv_direct(x, w, b):
    sx = 16.0 if x.abs().max() < 3000 else 1.0
    sw = pow2(w.abs().max().item(), 16384.0)
    xs, ws = x * sx, w * sw
    xh, wh = q16(xs), q16(ws)
    xl, wl = q16(xs - xh), q16(ws - wh)
    y = F.conv2d(xl, wh, None, padding=1) + F.conv2d(xh, wl, None, padding=1) + F.conv2d(xh, wh, None, padding=1)
    return y / (sx * sw) + b.view(1, -1, 1, 1)
def conv_winograd(x, w, b, name):
    B, C, H, W = x.shape
    Hp, Wp = (H + 1) // 2 * 2, (W + 1) // 2 * 2
    xp = F.pad(x, (1, 1 + Wp - W, 1, 1 + Hp - H))
    d = xp.unfold(2, 4, 2).unfold(3, 4, 2)                        # [B, C, th, tw, 4, 4]
    V = torch.einsum('ij,bcthjk,lk->bcthil', BT, d, BT)           # B^T d B, fp32
    U = torch.einsum('ij,ocjk,lk->ocil', G, w.double(), G).float()  # G g G^T (offline, fp64 -> fp32)
    sv = pow2(V.abs().max().item(), 2048.0)                       # transform-domain activations grow up to 4x
    su = pow2(U.abs().max().item(), 16384.0)
    st = STAT.setdefault(name, [0.0, 0.0])
    st[0] = max(st[0], x.abs().max().item()); st[1] = max(st[1], V.abs().max().item())
    Vs, Us = V * sv, U * su
    Vh, Uh = q16(Vs), q16(Us)
    Vl, Ul = q16(Vs - Vh), q16(Us - Uh)
    def gemm(a, u): return torch.einsum('bcthil,ocil->bothil', a, u)  # 16 independent channel contractions
    M = (gemm(Vl, Uh) + gemm(Vh, Ul) + gemm(Vh, Uh)) / (sv * su)
    Y = torch.einsum('ij,bothjk,lk->bothil', AT, M, AT)           # A^T M A -> [B, O, th, tw, 2, 2]
    th, tw = Y.shape[2], Y.shape[3]
    y = Y.permute(0, 1, 2, 4, 3, 5).reshape(B, -1, 2 * th, 2 * tw)[:, :, :H, :W]
    return y + b.view(1, -1, 1, 1)
def denoiser(x_t, t, cond, mode):
    P = 'depth_head.model.'
    def cv(x, name):
        w, b = sd[P + name + '.weight'], sd[P + name + '.bias']
        return conv_winograd(x, w, b, name) if (mode == 'winograd' and name in WINO) else conv_direct(x, w, b)
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
# single-layer check of the emulation itself
with torch.no_grad():
    xx = torch.randn(1, 256, 12, 20); ww = sd['depth_head.model.upsample_fuse.convA.conv.weight']; bb = sd['depth_head.model.upsample_fuse.convA.conv.bias']
    ref = F.conv2d(xx.double(), ww.double(), bb.double(), padding=1)
    e_w = (conv_winograd(xx, ww, bb, 'check').double() - ref).abs().max().item() / ref.abs().max().item()
    e_d = (conv_direct(xx, ww, bb).double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"one 256->256 layer vs fp64: winograd+split rel err {e_w:.2e}, direct 3-pass split {e_d:.2e}", flush=True)
for mode in ('winograd',):
    t0 = time.time()
    x = noise.clone()
    with torch.no_grad():
        for t in restate.ddim_timesteps(T):
            x = restate.ddim_step(denoiser(x, t, cond, mode), t, x, acp, T)
        z = restate.decode_logits(sd, x)
    dz = (dd_helpers.golden_view(g, 'logits', z) - zg).abs()
    print(f"{case} {mode}: {time.time() - t0:.0f} s  max|dz| vs reference golden {dz.max().item():.3e}  rms {dz.pow(2).mean().sqrt().item():.3e}", flush=True)
for k, v in STAT.items():
    print(f"  {k}: max|x| {v[0]:.1f}  max|B^T d B| {v[1]:.1f}")
