import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd
dev = torch.device('cuda:0')
eng = dd.DenoiseEngine('swin', 1, (8, 16), (4, 8), 2, dev, cuda_graph=False)
def tf(M, K, N, ms): return 2.0 * M * K * N / (ms * 1e-3) / 1e12
for (M, K, N, name) in [(6688, 768, 2304, 's2 qkv'), (6688, 768, 768, 's2 proj'), (6688, 768, 3072, 's2 ffn1'), (6688, 3072, 768, 's2 ffn2'),
                        (107008, 192, 576, 's0 qkv'), (107008, 192, 768, 's0 ffn1'), (107008, 768, 192, 's0 ffn2'),
                        (148 * 128 * 8, 768, 256, 'steady K768 N256 8 tiles/CTA'), (148 * 128 * 8, 3072, 256, 'steady K3072'),
                        (148 * 128, 768, 256, '1 tile/CTA K768'), (148 * 128, 3072, 256, '1 tile/CTA K3072'), (148*128, 96*32, 256, '1 tile K3072b')]:
    out = []
    for mode in (3, 0, 1, 2):
        ms = eng.bench_gemm(M, K, N, mode, 20)
        out.append(f"m{mode}: {ms*1e3:7.1f} us {tf(M,K,N,ms):6.0f} TF")
    print(f"{name:32s} M={M:7d} K={K:5d} N={N:5d} | " + " | ".join(out), flush=True)
