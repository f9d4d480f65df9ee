"""Experiment: small-batch rollouts are bound by per-kernel latency (about 7 us x 118 dependent kernels per model
pass).  Do K independent engines, each rolling out B/K clips on its own stream, overlap those latencies?"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sdvg_b200

cfg = sdvg_b200.CONFIGS["1_17_ball_complex_L1_64"]
B, C, P, W = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 10, 10, 5
dev = torch.device("cuda")
ctx = torch.randn(B, C, 256, device=dev)
for K in (1, 2, 4, 8):
    if B % K:
        continue
    models, streams, outs = [], [], []
    for k in range(K):
        torch.manual_seed(0)
        m = sdvg_b200.Transformer(0, cfg["dim_model"], cfg["num_heads"], cfg["num_encoder_layers"], cfg["num_decoder_layers"],
                                  0.1, frame_size=64, precision="mixed", max_clips=B // K, max_tokens=W, max_history=C + P).eval().to(dev)
        models.append(m); streams.append(torch.cuda.Stream()); outs.append(torch.empty(B // K, P, 256, device=dev))
    per = B // K
    pes = [sdvg_b200.pe_index_for(k * per, (k + 1) * per, dev) for k in range(K)]
    parts = [ctx[k * per:(k + 1) * per].contiguous() for k in range(K)]

    def step():
        for k in range(K):
            with torch.cuda.stream(streams[k]):
                sdvg_b200.rollout(models[k], parts[k], P, W, pe_index=pes[k], out=outs[k])
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"B={B} lanes={K}: {dt * 1e3:.2f} ms per rollout, {B * P / dt:.0f} frames/s", flush=True)
    del models
    torch.cuda.empty_cache()
