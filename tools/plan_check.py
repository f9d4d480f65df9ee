"""Is Engine::choose_plan's pick the fastest tile plan for a shape?  python tools/plan_check.py M [M ...]
Times sdvg_gemm (fp16, back-to-back launches) with block_n = 0 (the planner) and every explicit plan (negative = CTA pair)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdvg_b200

for M in [int(a) for a in sys.argv[1:]] or [640]:
    for (N, K) in ((2048, 2048), (6144, 2048)):
        g = torch.Generator(device="cuda").manual_seed(0)
        A = torch.randn(M, K, device="cuda", generator=g)
        W = torch.randn(N, K, device="cuda", generator=g) * 0.05
        row = []
        for bn in (0, -256, -192, -128, -64, 256, 128, 64, 32):
            try:
                _, ms = sdvg_b200.gemm(A, W, None, precision="fp16", block_n=bn, iters=200)
                row.append(f"{bn}: {ms * 1e3:.1f}")
            except Exception as e:
                row.append(f"{bn}: -")
        print(f"M={M} N={N} K={K} us/launch  " + "  ".join(row), flush=True)
