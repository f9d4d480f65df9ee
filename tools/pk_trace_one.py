"""Pipeline trace of one-op persistent-kernel programs: SDVG_PK_TRACE=1[,cta] python tools/pk_trace_one.py M N K [precision]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sdvg_b200
M, N, K = (int(v) for v in sys.argv[1:4])
prec = sys.argv[4] if len(sys.argv) > 4 else "fp16"
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.05; b = torch.randn(N, device="cuda")
X = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for it in range(3):
    for _ in range(60):          # ~100 ms of dense work right before: clocks are ramped when the traced launch runs
        X @ X
    C, ms = sdvg_b200.gemm(A, W, b, relu=True, precision=prec, block_n=9999, iters=1)
    print(f"{M}x{N}x{K} {prec}: {ms*1e3:.1f} us", flush=True)
