"""Every GEMM of the native Swin-L backbone at BASELINE config 3 (4 x 352 x 1216), timed alone with the epilogue it
runs with (dd_bench_gemm modes: 0 fp32 out, 1 fp32 out + residual, 2 GELU -> fp16 planes, 3 mainloop only)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd
dev = torch.device('cuda:0')
eng = dd.DenoiseEngine('swin', 1, (8, 16), (4, 8), 2, dev, cuda_graph=False)
B = 4
stages = [(88 * 304, 192, 2), (44 * 152, 384, 2), (22 * 76, 768, 18), (11 * 38, 1536, 2)]
total = ideal = 0.0
PEAK = 1500e12 / 3  # 3-pass split at ~1.5 PF issued
for s, (hw, C, depth) in enumerate(stages):
    M = B * hw
    rows = [("qkv", M, C, 3 * C, 0), ("proj", M, C, C, 1), ("ffn1", M, C, 4 * C, 2), ("ffn2", M, 4 * C, C, 1)]
    if s < 3: rows.append(("merge", M // 4, 4 * C, 2 * C, 0))
    for name, m, k, n, mode in rows:
        cnt = depth if name != "merge" else 1
        ms = eng.bench_gemm(m, k, n, mode, 10)
        ms3 = eng.bench_gemm(m, k, n, 3, 10)
        fl = 2.0 * m * k * n
        tiles = ((m + 127) // 128) * (n // (256 if n % 256 == 0 else 192))
        print(f"s{s} {name:5s} M={m:6d} K={k:5d} N={n:5d} x{cnt:2d}: {ms*1e3:7.1f} us ({fl/ms/1e9:5.0f} TF) mainloop-only {ms3*1e3:7.1f} us; "
              f"ideal {fl/PEAK*1e6:6.1f} us; tiles {tiles} = {tiles/148:.2f} waves", flush=True)
        total += cnt * ms; ideal += cnt * fl / PEAK * 1e3
print(f"sum of backbone GEMMs: {total:.2f} ms; at 500 TF algorithmic: {ideal:.2f} ms")
