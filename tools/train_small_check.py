"""Per-parameter gradient error of a small training step vs the oracle (debug aid): d H Le Ld B S P"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sdvg_b200
from oracle import train as OT
from oracle.ref_module import RefTransformer
KW = dict(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07, lambda_contrastive=0.1)
d, H, Le, Ld, B, S, P = (int(v) for v in sys.argv[1:8])
torch.manual_seed(int(os.environ.get("SEED", "12")))
ref = RefTransformer(0, d, H, Le, Ld, 0.0, frame_size=64)
m = sdvg_b200.Transformer(0, d, H, Le, Ld, 0.0, frame_size=64, precision="fp32")
m.load_state_dict(ref.state_dict()); m = m.to("cuda")
tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=P, **KW)
batch = OT.make_batch(B, S, 256, seed=int(os.environ.get("BATCH_SEED", "13")))
loss, pred, grads = OT.train_step_ref(ref, torch.optim.Adam(ref.parameters(), lr=1e-5), batch, P, **KW)
torch.manual_seed(int(os.environ.get("SEED", "12")))
r64 = RefTransformer(0, d, H, Le, Ld, 0.0, frame_size=64).double()
_, _, g64 = OT.train_step_ref(r64, torch.optim.Adam(r64.parameters(), lr=1e-5), batch.double(), P, **KW)
losses = tr.step(batch.to("cuda"))
print("loss", float(loss), float(losses[0]), "pred", float((tr.prediction(B, S - 1).cpu() - pred).abs().max() / pred.abs().max()))
rows = sorted(((float((tr.gradient(k).cpu() - g).abs().max() / g.abs().max()),
                float((tr.gradient(k).cpu().double() - g64[k]).abs().max() / g64[k].abs().max()),
                float((g.double() - g64[k]).abs().max() / g64[k].abs().max()), k) for k, g in grads.items()), reverse=True)
for e, e64, r, k in rows[:40]:
    print(f"  ours-ref32 {e:.2e}  ours-f64 {e64:.2e}  ref32-f64 {r:.2e}  {k}")
k = rows[0][3]
if k.endswith("linear1.weight"):
    kb = k.replace("weight", "bias")
    eb = (tr.gradient(kb).cpu().double() - g64[kb]).abs() / g64[kb].abs().max()
    ew = ((tr.gradient(k).cpu().double() - g64[k]).abs() / g64[k].abs().max()).max(dim=1).values
    print("bias entries off by > 1e-4:", int((eb > 1e-4).sum()), "at", (eb > 1e-4).nonzero().flatten().tolist()[:8],
          "| weight rows off by > 1e-4:", (ew > 1e-4).nonzero().flatten().tolist()[:8])
