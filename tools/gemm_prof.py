"""Micro-benchmark of one GEMM shape through the C ABI (used under ncu):  python tools/gemm_prof.py M N K prec bn iters"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdvg_b200

M, N, K = (int(v) for v in sys.argv[1:4])
prec = sys.argv[4] if len(sys.argv) > 4 else "fp16"
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, device="cuda", generator=g)
W = torch.randn(N, K, device="cuda", generator=g) * 0.05
_, ms = sdvg_b200.gemm(A, W, None, precision=prec, block_n=bn, iters=iters)
print(f"{prec} {M}x{N}x{K} bn={bn}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s")
