"""Profiling driver: one short-K GEMM of Swin stage 0 (qkv: M=107008, K=192, N=576, fp32 out) on the GEMM-mode conv kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusiondepth_b200 as dd
eng = dd.DenoiseEngine('swin', 1, (8, 16), (4, 8), 2, torch.device('cuda:0'), cuda_graph=False)
print(eng.bench_gemm(107008, 192, 576, 0, 1))
