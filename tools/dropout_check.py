"""Per-parameter gradient error of a dropout training step against the oracle under the same masks (debug aid)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sdvg_b200
from oracle import train as OT, dropout as D
from oracle.ref_module import RefTransformer
KW = dict(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07, lambda_contrastive=0.1)
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
seed = 0x123456789ABC
torch.manual_seed(5)
ref = RefTransformer(0, 64, 2, 2, 2, 0.0, frame_size=64)
m = sdvg_b200.Transformer(0, 64, 2, 2, 2, 0.0, frame_size=64, precision="fp32")
m.load_state_dict(ref.state_dict()); m = m.to("cuda")
tr = sdvg_b200.AdamTrainer(m, lr=1e-3, frames_to_predict=5, dropout=p, seed=seed, **KW)
sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
sd64 = {k: v.double() for k, v in sd.items()}
for step in (1, 2):
    batch = OT.make_batch(4, 6, 256, seed=30 + step)
    loss, pred, grads = OT.train_grads_functional(sd, 2, batch, 5, drop=D.Dropper(p, seed, step), **KW)
    l64, p64, g64 = OT.train_grads_functional(sd64, 2, batch.double(), 5, drop=D.Dropper(p, seed, step), **KW)
    losses = tr.step(batch.to("cuda"))
    print("step", step, "loss", float(loss), float(losses[0]), "pred", float((tr.prediction(4, 5).cpu() - pred).abs().max() / pred.abs().max()))
    rows = []
    for k, gr in grads.items():
        got = tr.gradient(k).cpu()
        sc = float(g64[k].abs().max())
        rows.append((float((got - gr).abs().max() / gr.abs().max()), float((got.double() - g64[k]).abs().max()) / sc,
                     float((gr.double() - g64[k]).abs().max()) / sc, k))
    rows.sort(reverse=True)
    for r in rows[:6]:
        print(f"  ours-ref32 {r[0]:.2e}  ours-f64 {r[1]:.2e}  ref32-f64 {r[2]:.2e}  {r[3]}")
    tr.pull_weights()
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    sd64 = {k: v.double() for k, v in sd.items()}
