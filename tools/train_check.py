"""Per-parameter gradient error of the training step against the oracle (debug aid; run on a B200)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import sdvg_b200
from oracle import train as OT
from oracle.ref_module import RefTransformer

name = sys.argv[1] if len(sys.argv) > 1 else "11_19_wallpushups_all_losses_test"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
c = sdvg_b200.CONFIGS[name]
kw = dict(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07, lambda_contrastive=0.1)
torch.manual_seed(0)
ref = RefTransformer(0, c["dim_model"], c["num_heads"], c["num_encoder_layers"], c["num_decoder_layers"], 0.0, frame_size=c["frame_size"])
m = sdvg_b200.Transformer(0, c["dim_model"], c["num_heads"], c["num_encoder_layers"], c["num_decoder_layers"], 0.0, frame_size=c["frame_size"], precision=prec)
m.load_state_dict(ref.state_dict()); m = m.to("cuda")
tr = sdvg_b200.AdamTrainer(m, lr=1e-5, frames_to_predict=5, **kw)
opt = torch.optim.Adam(ref.parameters(), lr=1e-5)
E = 4 * (c["frame_size"] // 8) ** 2
batch = OT.make_batch(B, 6, E, seed=9)
loss, pred, grads = OT.train_step_ref(ref, opt, batch, 5, **kw)
ref64 = None
if "--f64" in sys.argv:
    torch.manual_seed(0)
    r2 = RefTransformer(0, c["dim_model"], c["num_heads"], c["num_encoder_layers"], c["num_decoder_layers"], 0.0, frame_size=c["frame_size"]).double()
    o2 = torch.optim.Adam(r2.parameters(), lr=1e-5)
    _, _, ref64 = OT.train_step_ref(r2, o2, batch.double(), 5, **kw)
losses = tr.step(batch.to("cuda"))
print("loss", float(loss), losses.tolist())
p = tr.prediction(B, 5).cpu()
print("pred maxrel", float((p - pred).abs().max() / pred.abs().max()))
rows = []
for k, gr in grads.items():
    got = tr.gradient(k).cpu()
    e = float((got - gr).abs().max() / gr.abs().max())
    e64 = r64 = float("nan")
    if ref64 is not None:
        g64 = ref64[k]
        e64 = float((got.double() - g64).abs().max() / g64.abs().max())
        r64 = float((gr.double() - g64).abs().max() / g64.abs().max())
    rows.append((e, k, float(gr.abs().max()), e64, r64))
rows.sort(reverse=True)
for e, k, mx, e64, r64 in rows[:25]:
    print(f"{e:.3e}  ours-vs-f64 {e64:.3e}  ref32-vs-f64 {r64:.3e}  max|g| {mx:.3e}  {k}")
