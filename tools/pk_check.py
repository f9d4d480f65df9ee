"""First-light / regression driver for the persistent small-batch kernel (csrc/persistent.cuh).

    python tools/pk_check.py gemm       # one-op programs through sdvg_gemm(block_n=9999) vs float64
    python tools/pk_check.py forward    # small model: persistent pass vs per-kernel pass vs oracle
    python tools/pk_check.py rollout    # C1 (d2048 4e/8d, B=8): parity vs per-kernel path + ms per pass
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sdvg_b200  # noqa: E402


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def ref(A, W, bias, relu, precision):
    if precision in ("fp16", "bf16"):
        dt = torch.float16 if precision == "fp16" else torch.bfloat16
        A, W = A.to(dt), W.to(dt)
    y = A.double() @ W.double().t()
    if bias is not None:
        y = y + bias.double()
    return torch.relu(y) if relu else y


def gemm():
    shapes = [(40, 128, 512), (40, 2048, 2048), (16, 256, 64), (8, 256, 2048), (48, 2048, 256), (80, 6144, 2048),
              (96, 1024, 1024), (128, 512, 1024), (5, 200, 512), (77, 96, 32), (1, 32, 8), (40, 32768, 2048)]
    worst = 0.0
    for precision in ("fp16", "fp32", "bf16"):
        for (M, N, K) in shapes:
            g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
            A = torch.randn(M, K, device="cuda", generator=g)
            W = torch.randn(N, K, device="cuda", generator=g) * 0.05
            b = torch.randn(N, device="cuda", generator=g)
            C1, ms = sdvg_b200.gemm(A, W, b, relu=True, precision=precision, block_n=9999, iters=1)
            C2, ms = sdvg_b200.gemm(A, W, b, relu=True, precision=precision, block_n=9999, iters=300)
            e = relerr(C1, ref(A, W, b, True, precision))
            worst = max(worst, e)
            bytes_w = N * ((K + 63) // 64 * 64) * 2 * (2 if precision == "fp32" else 1)
            print(f"pk gemm {precision} {M}x{N}x{K}: max-rel {e:.2e} deterministic {torch.equal(C1, C2)} "
                  f"{ms * 1e3:.1f} us/launch {bytes_w / (ms * 1e-3) / 1e9:.0f} GB/s", flush=True)
    print("worst", worst)
    assert worst < 2e-5


def build(d, H, Le, Ld, precision, pk, fs=64, **kw):
    os.environ["SDVG_PK"] = "1" if pk else "0"
    torch.manual_seed(0)
    m = sdvg_b200.Transformer(0, d, H, Le, Ld, 0.1, frame_size=fs, precision=precision, **kw).eval().cuda()
    m.engine(torch.device("cuda", 0))      # SDVG_PK is read when the engine is created
    return m


def forward():
    from oracle import functional as F
    for (d, H, Le, Ld) in ((128, 4, 2, 2), (256, 8, 1, 2), (512, 8, 2, 1)):
        for precision in ("fp32", "fp16", "mixed", "bf16"):
            mp, mc = build(d, H, Le, Ld, precision, True), build(d, H, Le, Ld, precision, False)
            sd = {k: v.detach().cpu() for k, v in mc.state_dict().items()}
            g = torch.Generator().manual_seed(1234)
            src = torch.randn(6, 6, 256, generator=g)
            tgt = src[:, :-1].contiguous()
            want = F.forward(sd, src, tgt, H, F.causal_mask(5))
            a = mc(src.cuda(), tgt.cuda(), mc.get_tgt_mask(5)).cpu()      # per-kernel path
            n0 = mp.launch_count()
            b = mp(src.cuda(), tgt.cuda(), mp.get_tgt_mask(5)).cpu()      # persistent path
            n1 = mp.launch_count()
            c = mp(src.cuda(), src.cuda(), "causal").cpu()
            c2 = mc(src.cuda(), src.cuda(), "causal").cpu()
            print(f"forward d{d} {precision}: per-kernel vs oracle {relerr(a, want):.2e}, persistent vs oracle {relerr(b, want):.2e}, "
                  f"persistent vs per-kernel {relerr(b, a):.2e} / same-src {relerr(c, c2):.2e}; launches of the persistent call {n1 - n0}",
                  flush=True)


def rollout():
    cfg = sdvg_b200.CONFIGS["1_17_ball_complex_L1_64"]
    B, C, P = 8, 10, 10
    g = torch.Generator().manual_seed(1234)
    ctx = torch.randn(B, C, 256, generator=g).cuda()
    for precision in ("mixed", "fp32"):
        outs = {}
        for pk in (False, True):
            m = build(cfg["dim_model"], cfg["num_heads"], cfg["num_encoder_layers"], cfg["num_decoder_layers"], precision, pk,
                      max_clips=B, max_tokens=10, max_history=C + P)
            for W in (5, 10):
                out = m.rollout(ctx, P, W)
                for _ in range(3):
                    m.rollout(ctx, P, W, out=out)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n0 = m.launch_count()
                e0.record()
                for _ in range(10):
                    m.rollout(ctx, P, W, out=out)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                outs[(pk, W)] = out.clone()
                print(f"rollout C1 {precision} W={W} persistent={pk}: {ms:.3f} ms per rollout, {ms / P * 1e3:.1f} us per pass, "
                      f"{B * P / ms * 1e3:.0f} frames/s, launches per rollout {(m.launch_count() - n0) / 10:.0f}", flush=True)
            del m
        for W in (5, 10):
            from oracle import rollout as R
            err = R.max_rel_per_frame(outs[(True, W)].cpu(), outs[(False, W)].cpu())
            print(f"  {precision} W={W}: persistent vs per-kernel per-frame max-rel {[f'{float(x):.1e}' for x in err]}", flush=True)


if __name__ == "__main__":
    t0 = time.time()
    for what in sys.argv[1:] or ["gemm", "forward", "rollout"]:
        {"gemm": gemm, "forward": forward, "rollout": rollout}[what]()
    print(f"done in {time.time() - t0:.1f} s")
