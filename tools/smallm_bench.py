"""Small-M GEMM timing: split-K plans against the single-CTA-per-tile plan (run on a B200)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sdvg_b200

for prec in ("fp32", "fp16"):
    for (M, N, K) in ((96, 1024, 1024), (96, 3072, 1024), (96, 2048, 1024), (96, 1024, 2048), (40, 2048, 2048), (40, 6144, 2048), (8, 2048, 2048)):
        A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.05
        row = []
        for bn, ks in ((32, 1), (64, 1), (128, 1), (32, 2), (32, 4), (64, 2), (64, 4), (128, 4), (128, 8), (0, 0)):
            if bn and (-(-N // bn)) * ks > 148:
                row.append(f"{bn}x{ks}: -")
                continue
            try:
                _, ms = sdvg_b200.gemm(A, W, None, precision=prec, block_n=(1000 * ks + bn) if ks > 1 else bn, iters=50)
                row.append(f"{bn}x{ks}: {ms * 1e3:.1f}")
            except RuntimeError as e:
                row.append(f"{bn}x{ks}: err")
        print(prec, (M, N, K), "us:", "  ".join(row), flush=True)
