"""Diagnostic for tests/test_gpu_train.py::test_c5_bench_shape_step_with_dropout_vs_same_mask_oracle: prints, per batch,
the distance of libsdvg and of the fp32 oracle from the float64 oracle (prediction, loss, worst gradient tensors)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import sdvg_b200
from oracle import train as OT
from oracle import dropout as D
from test_gpu_train import build_pair, CASES

c = sdvg_b200.CONFIGS["11_19_wallpushups_all_losses_test"]
arch = (c["dim_model"], c["num_heads"], c["num_encoder_layers"], c["num_decoder_layers"])
lr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
m, ref = build_pair(*arch, seed=0, frame_size=c["frame_size"])
m.dropout_p = 0.1
seed = 0x5EED_0C5
sd = ref.state_dict()
tr = sdvg_b200.AdamTrainer(m, lr=lr, frames_to_predict=5, seed=seed, **CASES["c5"])
for step, bseed in enumerate((12, 13, 14), start=1):
    batch = OT.make_batch(16, 6, 1024, seed=bseed)
    l32, p32, g32 = OT.train_grads_functional(sd, arch[1], batch, 5, drop=D.Dropper(0.1, seed, step), **CASES["c5"])
    l64, p64, g64 = OT.train_grads_functional({k: v.double() for k, v in sd.items()}, arch[1], batch.double(), 5,
                                              drop=D.Dropper(0.1, seed, step), **CASES["c5"])
    losses = tr.step(batch.to("cuda"))
    pred = tr.prediction(16, 5).cpu().double()
    e_pred = float((pred - p64).abs().max() / p64.abs().max())
    e_pred32 = float((p32.double() - p64).abs().max() / p64.abs().max())
    print(f"batch {bseed} step {step}: loss ours {float(losses[0]):.7f} f32 {float(l32):.7f} f64 {float(l64):.7f} | pred ours {e_pred:.2e} f32 {e_pred32:.2e}")
    d = (pred - p64).abs() / p64.abs().max()
    print("   pred rows with error > 1e-4:", sorted(set(map(tuple, (d > 1e-4).nonzero()[:, :2].tolist())))[:20], " elements:", int((d > 1e-4).sum()), "of", d.numel())
    rows = []
    for k, gr in g64.items():
        s = float(gr.abs().max())
        ours = float((tr.gradient(k).cpu().double() - gr).abs().max()) / s
        r32 = float((g32[k].double() - gr).abs().max()) / s
        fro = float((tr.gradient(k).cpu().double() - gr).norm() / gr.norm())
        fro32 = float((g32[k].double() - gr).norm() / gr.norm())
        rows.append((ours, r32, fro, fro32, k))
    rows.sort(reverse=True)
    for r in rows[:8]:
        print("   %-55s max-rel ours %.2e f32 %.2e | fro ours %.2e f32 %.2e" % (r[4], r[0], r[1], r[2], r[3]))
    # kink signature: a flipped ReLU unit j of layer L changes ROW j of that layer's linear1.weight gradient by O(1)
    for r in rows[:40]:
        if r[4].endswith("linear1.weight") and r[0] > 1e-3:
            e = (tr.gradient(r[4]).cpu().double() - g64[r[4]]).abs() / g64[r[4]].abs().max()
            per_row = e.max(dim=1).values
            top = torch.topk(per_row, 4)
            print("      %s: rows with max error: %s -> %s ; median row error %.2e" % (r[4], top.indices.tolist(), ["%.1e" % v for v in top.values.tolist()], float(per_row.median())))
    strict = [k for k in g64 if k.startswith("out.") or k.startswith("transformer.decoder.norm.") or k.startswith("transformer.decoder.layers.11.norm3") or k.startswith("transformer.decoder.layers.11.linear2")]
    for r in rows:
        if r[4] in strict:
            print("   kink-free %-45s max-rel ours %.2e f32 %.2e" % (r[4], r[0], r[1]))
    import statistics
    print("   median max-rel ours %.2e f32 %.2e; tensors over 1e-4: ours %d f32 %d of %d" % (
        statistics.median(r[0] for r in rows), statistics.median(r[1] for r in rows),
        sum(r[0] > 1e-4 for r in rows), sum(r[1] > 1e-4 for r in rows), len(rows)))
