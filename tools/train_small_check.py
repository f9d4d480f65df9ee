"""Per-parameter gradient error of a small training step vs the oracle (debug aid): d H Le Ld B S P"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sdvg_b200
from oracle import train as OT
from oracle.ref_module import RefTransformer
KW = dict(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07, lambda_contrastive=0.1)
d, H, Le, Ld, B, S, P = (int(v) for v in sys.argv[1:8])
torch.manual_seed(12)
ref = RefTransformer(0, d, H, Le, Ld, 0.0, frame_size=64)
m = sdvg_b200.Transformer(0, d, H, Le, Ld, 0.0, frame_size=64, precision="fp32")
m.load_state_dict(ref.state_dict()); m = m.to("cuda")
tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=P, **KW)
batch = OT.make_batch(B, S, 256, seed=13)
loss, pred, grads = OT.train_step_ref(ref, torch.optim.Adam(ref.parameters(), lr=1e-5), batch, P, **KW)
losses = tr.step(batch.to("cuda"))
print("loss", float(loss), float(losses[0]), "pred", float((tr.prediction(B, S - 1).cpu() - pred).abs().max() / pred.abs().max()))
rows = sorted(((float((tr.gradient(k).cpu() - g).abs().max() / g.abs().max()), k) for k, g in grads.items()), reverse=True)
for e, k in rows[:40]:
    print(f"  {e:.2e}  {k}")
