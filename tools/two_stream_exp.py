"""Experiment: do two independent half-batch rollouts on two CUDA streams overlap each other's non-GEMM phases?
    python tools/two_stream_exp.py [clips_total]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdvg_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = sdvg_b200.CONFIGS["1_15_kitti_L1_64"]
dev = torch.device("cuda:0")


def make(b):
    torch.manual_seed(0)
    m = sdvg_b200.Transformer(0, cfg["dim_model"], cfg["num_heads"], cfg["num_encoder_layers"], cfg["num_decoder_layers"],
                              0.1, frame_size=64, precision="mixed", max_clips=b, max_tokens=5, max_history=20)
    return m.eval().to(dev)


def bench(models, ctxs, outs, streams, iters=6):
    def once():
        for m, c, o, s in zip(models, ctxs, outs, streams):
            with torch.cuda.stream(s):
                sdvg_b200.rollout(m, c, 10, 5, out=o)
    for _ in range(4):
        once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    return dt


g = torch.Generator().manual_seed(1)
ctx = torch.randn(B, 10, 256, generator=g).to(dev)
one = make(B)
dt1 = bench([one], [ctx], [torch.empty(B, 10, 256, device=dev)], [torch.cuda.Stream()])
print(f"one handle, {B} clips: {dt1*1e3:.2f} ms -> {B*10/dt1:.0f} frames/s", flush=True)
del one
torch.cuda.empty_cache()
for parts in (2, 3, 4):
    b = B // parts
    ms = [make(b) for _ in range(parts)]
    cs = [ctx[i * b:(i + 1) * b].contiguous() for i in range(parts)]
    os_ = [torch.empty(b, 10, 256, device=dev) for _ in range(parts)]
    ss = [torch.cuda.Stream() for _ in range(parts)]
    dt = bench(ms, cs, os_, ss)
    print(f"{parts} handles x {b} clips on {parts} streams: {dt*1e3:.2f} ms -> {b*parts*10/dt:.0f} frames/s", flush=True)
    del ms
    torch.cuda.empty_cache()
