"""Host enqueue time vs device time of the training step (is the step launch-bound on the CPU?)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sdvg_b200
c = sdvg_b200.CONFIGS["11_19_wallpushups_all_losses_test"]
torch.manual_seed(0)
m = sdvg_b200.Transformer(0, c["dim_model"], c["num_heads"], c["num_encoder_layers"], c["num_decoder_layers"], 0.0,
                          frame_size=c["frame_size"], precision="fp32").to("cuda")
tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=5)
x = torch.randn(16, 6, 1024, device="cuda")
for _ in range(3):
    tr.step(x)
torch.cuda.synchronize()
n = 20
t0 = time.perf_counter()
for _ in range(n):
    tr.step(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / n:.2f} ms/step, total {1e3 * (t2 - t0) / n:.2f} ms/step, launches/step {m.launch_count() // 23}")
hs, ts = [], []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); tr.step(x); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    hs.append(1e3 * (t1 - t0)); ts.append(1e3 * (t2 - t0))
print("single step from an idle queue: host", [round(v, 2) for v in hs], "total", [round(v, 2) for v in ts])
