"""GPU bring-up checks (run under gpurun).  Each check runs in its own subprocess with a timeout so a
device-side trap in one kernel cannot poison the others.  Output: gpurun_out/check.log
    python tools/gpu_check.py            # all checks
    python tools/gpu_check.py gemm_tc    # one check (in-process)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def ref_gemm(A, W, bias, relu, precision):
    import torch
    if precision in ("fp16", "bf16"):
        dt = torch.float16 if precision == "fp16" else torch.bfloat16
        A, W = A.to(dt), W.to(dt)
    y = A.double() @ W.double().t()
    if bias is not None:
        y = y + bias.double()
    return torch.relu(y) if relu else y


def check_gemm_simt():
    import torch, sdvg_b200
    g = torch.Generator(device="cuda").manual_seed(0)
    for (M, N, K) in ((200, 96, 64), (130, 300, 256), (40, 2048, 2048)):
        A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(N, K, device="cuda", generator=g) * 0.05
        b = torch.randn(N, device="cuda", generator=g)
        C, _ = sdvg_b200.gemm(A, W, b, relu=True, precision="fp32_simt")
        print(f"simt {M}x{N}x{K}: relerr {relerr(C, ref_gemm(A, W, b, True, 'fp32')):.2e}", flush=True)


def check_gemm_pattern():
    """Layout probe: A rows are unit vectors, so C[i,n] = W[n, i % K] exactly in any precision."""
    import torch, sdvg_b200
    for prec in ("fp16", "fp32"):
        for bn in (32, 64, 128, 256, -64, -128, -192, -256):
            if prec == "fp32" and abs(bn) > 128:
                continue
            M, N, K = (128, 256, 64) if bn > 0 else (512, 384, 128)
            A = torch.zeros(M, K, device="cuda"); A[torch.arange(M), torch.arange(M) % K] = 1.0
            W = (torch.arange(N * K, device="cuda", dtype=torch.float32).view(N, K) % 251) / 16.0
            C, _ = sdvg_b200.gemm(A, W, None, precision=prec, block_n=bn)
            want = W.t()[torch.arange(M) % K]
            err = relerr(C, want)
            print(f"pattern {prec} bn={bn}: relerr {err:.2e}", flush=True)
            if err > 1e-3:
                bad = (C - want).abs() > 1e-3
                print("  first bad rows/cols:", bad.nonzero()[:8].tolist())
                print("  C[0:4,0:8]   ", C[0:4, 0:8].tolist())
                print("  want[0:4,0:8]", want[0:4, 0:8].tolist())


def check_gemm_tc():
    import torch, sdvg_b200
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = ((128, 256, 64), (128, 256, 256), (300, 520, 2048), (77, 96, 32), (1000, 6144, 2048), (5120, 256, 2048))
    for prec in ("fp16", "bf16", "fp32"):
        for (M, N, K) in shapes:
            A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(N, K, device="cuda", generator=g) * 0.05
            b = torch.randn(N, device="cuda", generator=g)
            for bn in (0, 32, 64, 128, 256, -64, -128, -192, -256):
                if (prec == "fp32" and abs(bn) > 128) or (bn > N and bn > 32) or (-bn // 2 > N):
                    continue
                C, _ = sdvg_b200.gemm(A, W, b, relu=False, precision=prec, block_n=bn)
                e = relerr(C, ref_gemm(A, W, b, False, prec))
                print(f"tc {prec} {M}x{N}x{K} bn={bn}: relerr {e:.2e} {'OK' if e < (2e-6 if prec=='fp32' else 2e-5) else 'BAD'}", flush=True)


def check_gemm_speed():
    import torch, sdvg_b200
    g = torch.Generator(device="cuda").manual_seed(0)
    for (M, N, K) in ((5120, 6144, 2048), (5120, 2048, 2048), (5120, 4096, 2048), (10240, 2048, 2048), (8192, 8192, 8192), (5120, 2048, 256), (5120, 256, 2048), (40, 6144, 2048), (80, 2048, 2048)):
        A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(N, K, device="cuda", generator=g) * 0.05
        for prec in ("fp16", "fp32"):
            for bn in (64, 128, 256, -64, -128, -192, -256, 0):
                if prec == "fp32" and abs(bn) > 128:
                    continue
                sdvg_b200.gemm(A, W, None, precision=prec, block_n=bn, iters=2)
                _, ms = sdvg_b200.gemm(A, W, None, precision=prec, block_n=bn, iters=10)
                print(f"speed {prec} {M}x{N}x{K} bn={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
        a16, w16 = A.half(), W.half()
        for _ in range(3): a16 @ w16.t()
        torch.cuda.synchronize(); t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10): a16 @ w16.t()
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print(f"speed cublas-fp16 {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)


def _model_from_golden(g, precision):
    import torch, sdvg_b200
    from conftest import ref_model_from_golden
    ref = ref_model_from_golden(g)
    d, H, Le, Ld, E = (int(v) for v in g["arch"])
    m = sdvg_b200.Transformer(0, d, H, Le, Ld, 0.1, frame_size={256: 64, 1024: 128}[E], precision=precision)
    m.load_state_dict(ref.state_dict())
    return m.eval().cuda(), ref


def check_forward(precisions=("fp32_simt", "fp32", "fp16", "mixed", "bf16")):
    import torch
    from conftest import load_golden
    g = load_golden("tiny_forward")
    for prec in precisions:
        m, ref = _model_from_golden(g, prec)
        with torch.no_grad():
            for key, (s, t, mask) in {"out_causal": (g["src"], g["tgt"], ref.get_tgt_mask(5)),
                                      "out_nomask": (g["src"], g["tgt"], None),
                                      "out_same": (g["src"], g["src"], "causal"),
                                      "out_b64": (g["x64"], g["x64"], ref.get_tgt_mask(2))}.items():
                s, t = s.cuda(), t.cuda()
                o = m(s, s if key in ("out_same", "out_b64") else t, mask)
                print(f"forward tiny {prec} {key}: relerr {relerr(o.cpu(), g[key]):.2e}", flush=True)
    g = load_golden("d256_forward")
    for prec in precisions:
        m, ref = _model_from_golden(g, prec)
        with torch.no_grad():
            x = g["x"].cuda()
            print(f"forward d256 {prec}: relerr {relerr(m(x, x, 'causal').cpu(), g['out']):.2e}", flush=True)


def check_rollout_small(precisions=("fp32_simt", "fp32", "fp16", "mixed")):
    import torch, sdvg_b200
    from conftest import load_golden
    from oracle import rollout as R
    g = load_golden("small_rollout")
    for prec in precisions:
        m, ref = _model_from_golden(g, prec)
        ctx = g["ctx"].cuda()
        f5 = sdvg_b200.rollout(m, ctx, 4, 5).cpu()
        f10 = sdvg_b200.rollout(m, ctx, 3, 10).cpu()
        fa = sdvg_b200.rollout(m, g["frames"].cuda(), 4, 5, use_sos=True).cpu()
        tf = sdvg_b200.rollout(m, ctx, 4, 5, teacher=g["free5"].cuda()).cpu()
        print(f"rollout small {prec}: free5 {R.max_rel_per_frame(f5, g['free5']).tolist()}", flush=True)
        print(f"rollout small {prec}: free10 {R.max_rel_per_frame(f10, g['free10']).tolist()}", flush=True)
        print(f"rollout small {prec}: faithful {R.max_rel_per_frame(fa, g['faithful']).tolist()}", flush=True)
        print(f"rollout small {prec}: teacher {R.max_rel_per_frame(tf, g['free5']).tolist()}", flush=True)


def check_rollout_c1(precisions=("fp32", "fp16", "mixed", "bf16", "fp32_simt")):
    import torch, sdvg_b200
    from conftest import load_golden
    from oracle import rollout as R
    for name in ("c1_rollout", "c4_rollout"):
        g = load_golden(name)
        for prec in precisions:
            m, ref = _model_from_golden(g, prec)
            ctx = g["ctx"].cuda()
            n = g["free5"].shape[1]
            fr = sdvg_b200.rollout(m, ctx, n, 5).cpu()
            tf = sdvg_b200.rollout(m, ctx, n, 5, teacher=g["free5"].cuda()).cpu()
            print(f"rollout {name} {prec}: free {['%.2e' % v for v in R.max_rel_per_frame(fr, g['free5'])]} "
                  f"teacher {['%.2e' % v for v in R.max_rel_per_frame(tf, g['free5'])]}", flush=True)
            del m
            torch.cuda.empty_cache()


CHECKS = {k[6:]: v for k, v in list(globals().items()) if k.startswith("check_")}

if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--run":
        CHECKS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(CHECKS)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "check.log"), "a")
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", name], capture_output=True, text=True, timeout=600)
            out, rc = r.stdout + r.stderr[-3000:], r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else str(e.stdout), "TIMEOUT"
        msg = f"===== {name}: rc={rc} ({time.time()-t0:.1f}s)\n{out}\n"
        log.write(msg); log.flush()
        print(msg, flush=True)
