"""bench_train.py - the data-parallel training step (BASELINE.json configs[4]) on N B200s.

    python bench_train.py [--gpus N --steps K --warmup W]          # libsdvg (sm_100a kernels), one rank per GPU
    python bench_train.py --impl reference [...]                    # the reference's step on the host CPU (oracle port)
    torchrun --nproc-per-node N ... bench_train.py --gpus N ...     # N > 1: NCCL gradient all-reduce

A "step" is one iteration of Trainer.train_loop (trainers/trainer.py:123-162) on config
11_19_wallpushups_all_losses_test: d1024 H16 12enc/12dec, E=1024, 16 clips per GPU of 6 latents (SOS + 5 frames),
loss = MSE + GDL(alpha 2) + 0.1 BiPatchNCE, Adam lr 1e-5, dropout 0.1 (DROPOUT_P of the config; ours draws its masks
from the library's counter-based generator, the CPU arm from torch's); synthetic latents, seeded random-init weights.  Weak scaling: the global batch is 16 N clips, gradients are averaged over
ranks (NCCL all-reduce of the flat fp32 gradient vector in two buckets, the first overlapped with the encoder
backward).  value = clips/s with the batch resident in HBM; e2e = the same with the batch copied from pinned host
memory and the loss read back every step.  The step is HBM-bound (M = 80..96 rows per GEMM): the roofline is the
algorithmic parameter traffic per step against the measured HBM bandwidth."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LOSS = dict(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07,
            lambda_contrastive=0.1)


def parse(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", default="11_19_wallpushups_all_losses_test")
    p.add_argument("--batch", type=int, default=16, help="clips per GPU (BATCH_SIZE of the config)")
    p.add_argument("--precision", default="fp32", choices=["fp32", "fp16", "bf16", "mixed"])
    p.add_argument("--no-overlap", action="store_true")
    p.add_argument("--no-dropout", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args(argv)


def param_count(cfg, E):
    d, ff, Le, Ld = cfg["dim_model"], 2048, cfg["num_encoder_layers"], cfg["num_decoder_layers"]
    attn = 4 * d * d + 4 * d
    ffn = 2 * d * ff + ff + d
    return E * d + d + Le * (attn + ffn + 4 * d) + 2 * d + Ld * (2 * attn + ffn + 6 * d) + 2 * d + d * E + E


def cpu_step(cfg, E, batch, steps):
    import torch
    from oracle import train as OT
    from oracle.ref_module import RefTransformer
    torch.manual_seed(0)
    ref = RefTransformer(0, cfg["dim_model"], cfg["num_heads"], cfg["num_encoder_layers"], cfg["num_decoder_layers"],
                         cfg["dropout_p"], frame_size=cfg["frame_size"])
    opt = torch.optim.Adam(ref.parameters(), lr=1e-5)
    times = []
    for i in range(steps + 1):
        x = OT.make_batch(batch, 6, E, seed=50 + i)
        t0 = time.perf_counter()
        OT.train_step_ref(ref, opt, x, 5, **LOSS)
        times.append(time.perf_counter() - t0)
    return sum(times[1:]) / steps


def main(argv=None):
    out = measure(parse(argv))
    if out is not None:
        print(json.dumps(out))


def measure(a, quick=False):
    """Runs the benchmark and returns the JSON line as a dict on rank 0 (None elsewhere).  `quick` (bench.py's `extra`
    block): no end-to-end leg, no CPU baseline, the process group of the caller is reused."""
    import torch
    import sdvg_b200
    cfg = sdvg_b200.CONFIGS[a.config]
    E = sdvg_b200.latent_dim(cfg["frame_size"])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric, unit = "train_clips_per_sec", "clips/s"
    workload = (f"{a.config} training step: {a.batch} clips/GPU x {max(world, 1)} GPU, S_src 6 / S_tgt 5, "
                "MSE+GDL(2)+0.1 BiPatchNCE, Adam, dropout " + ("0" if a.no_dropout else str(cfg["dropout_p"])))
    base = {"metric": metric, "unit": unit, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "data": "synthetic", "config": {"workload": workload, "arch": cfg, "clips_per_gpu": a.batch,
                                                     "dropout": 0.0 if a.no_dropout else cfg["dropout_p"]}}
    if a.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        steps = max(1, min(a.steps, 5))
        dt = cpu_step(cfg, E, a.batch, steps)
        v = a.batch / dt
        return {**base, "impl": "reference", "value": v, "n_gpus": 0, "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3,
                "dtype": "f32", "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port",
                                                 "sample": f"{steps} full steps of {a.batch} clips"},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)      # same weights on every rank (and as the CPU arm: the module initialises like the reference)
    m = sdvg_b200.Transformer(0, cfg["dim_model"], cfg["num_heads"], cfg["num_encoder_layers"], cfg["num_decoder_layers"],
                              cfg["dropout_p"], frame_size=cfg["frame_size"], precision=a.precision)
    m = m.to(dev)
    tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=5, overlap=not a.no_overlap, seed=1234 + rank,
                               dropout=0.0 if a.no_dropout else None, **LOSS)
    def make_batch(seed):     # latents like encode_batch(..., use_sos=True): SOS frame of 2.0 + 5 unit-variance frames
        x = torch.randn(a.batch, 6, E, generator=torch.Generator().manual_seed(seed))
        x[:, 0] = 2.0
        return x
    host = [make_batch(1000 + rank * 131 + i).pin_memory() for i in range(4)]
    devb = [h.to(dev) for h in host]
    pe = (torch.arange(rank * a.batch, (rank + 1) * a.batch) % 64).to(dev, torch.int32)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    for i in range(max(3, a.warmup)):
        tr.step(devb[i % 4], pe_index=pe)
    from bench import ClockSampler                       # nvidia-smi clocks / throttle reasons during the timed region
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = m.launch_count()
    ms = timed(lambda i: tr.step(devb[i % 4], pe_index=pe), a.steps)
    launches = (m.launch_count() - n0) // a.steps

    loss_host = torch.empty(5).pin_memory()

    def e2e_step(i):
        x = host[i % 4].to(dev, non_blocking=True)
        loss_host.copy_(tr.step(x, pe_index=pe), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_e2e = ms if quick else timed(e2e_step, a.steps)

    clocks = sampler.stop() if sampler else None
    # per-kernel-class device time of one step (instrumented pass, outside the timed regions)
    m.timing(True)
    tr.step(devb[0], pe_index=pe)
    classes = {k: round(v["ms"], 3) for k, v in m.timing_read().items()}
    gemm_flops = None
    m.timing(False)
    del tr
    m._free()
    if rank != 0:
        return None
    P = param_count(cfg, E)
    split = a.precision == "fp32"
    plane = 4.0 if split else 2.0
    # algorithmic HBM bytes per step: forward W planes + backward W^T planes read, fp32 gradients written, Adam
    # (read p, g, m, v; write p, m, v) and the rebuild of both plane sets (read p, write 2 plane sets)
    bytes_step = P * (plane + plane + 4.0 + 28.0 + 4.0 + 2 * plane)
    hbm = 6545.9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    src = "fallback"
    if os.path.exists(peaks_path):
        hbm = json.load(open(peaks_path)).get("hbm_gbs", hbm)
        src = "measured (MEASURED_PEAKS.json)"
    out = {**base, "value": a.batch * world / (ms * 1e-3), "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
           "ms_per_step": ms, "dtype": "fp16x2 split operands, fp32 accumulate" if split else a.precision,
           "e2e": {"value": a.batch * world / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": host[0].numel() * 4,
                   "d2h_bytes_per_step": 20, "ms_per_step": ms_e2e},
           "gpu_launches": int(launches) * a.steps, "clocks": clocks,
           "roofline": {"bound": "hbm", "achieved": bytes_step / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                        "frac": bytes_step / (ms * 1e-3) / 1e9 / hbm, "traffic": None, "peak_source": src,
                        "algorithmic_bytes_per_step": bytes_step, "launches_per_step": int(launches), "classes_ms": classes}}
    if not a.no_cpu_baseline and not quick:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        dt = cpu_step(cfg, E, a.batch, 2)
        out["cpu_baseline"] = {"value": a.batch / dt, "unit": unit, "cores": cores, "kind": "port",
                               "sample": f"2 full steps of {a.batch} clips (oracle/train.py)", "ms_per_step": dt * 1e3}
    return out


if __name__ == "__main__":
    main()
