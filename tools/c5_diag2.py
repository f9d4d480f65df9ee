"""Part (b) of test_c5_bench_shape_step_with_dropout_vs_same_mask_oracle with a per-tensor report: backward pass through the
autograd bridge from the float64 oracle's dL/dpred; which tensors carry the whole-vector error, and is it one row?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import sdvg_b200
from oracle import train as OT
from oracle import dropout as D
from test_gpu_train import build_pair, CASES

c = sdvg_b200.CONFIGS["11_19_wallpushups_all_losses_test"]
arch = (c["dim_model"], c["num_heads"], c["num_encoder_layers"], c["num_decoder_layers"])
m, ref = build_pair(*arch, seed=0, frame_size=c["frame_size"])
m.dropout_p = 0.1
seed = 0x5EED_0C5
sd64 = {k: v.double() for k, v in ref.state_dict().items()}
m.train(); m.dropout_seed = seed
for step, bseed in enumerate((int(a) for a in (sys.argv[1:] or ["14", "12"])), start=1):
    batch = OT.make_batch(16, 6, 1024, seed=bseed)
    _, pred64, g64, dpred64 = OT.train_grads_functional(sd64, arch[1], batch.double(), 5, drop=D.Dropper(0.1, seed, step),
                                                         return_dpred=True, **CASES["c5"])
    x = batch.cuda()
    pred = m(x, x[:, :-1].contiguous(), m.get_tgt_mask(5).cuda())
    print(f"batch {bseed} step {step}: pred max-rel {float((pred.detach().cpu().double() - pred64).abs().max() / pred64.abs().max()):.2e}")
    m.zero_grad()
    pred.backward(dpred64.float().cuda())
    params = dict(m.named_parameters())
    rows, num, den = [], 0.0, 0.0
    for k, gr in g64.items():
        d = params[k].grad.cpu().double() - gr
        n2, g2 = float(d.pow(2).sum()), float(gr.pow(2).sum())
        num += n2; den += g2
        rows.append((n2, k, float(d.abs().max() / gr.abs().max()), (n2 / g2) ** 0.5, g2, d, gr))
    print(f"   whole-vector relative error {(num / den) ** 0.5:.3e}")
    rows.sort(key=lambda r: -r[0])
    for n2, k, mx, fro, g2, d, gr in rows[:8]:
        line = f"   {k:55s} share of error energy {n2 / num:6.1%}  share of gradient energy {g2 / den:6.1%}  max-rel {mx:.2e}  fro {fro:.2e}"
        if d.dim() == 2:
            pr = d.pow(2).sum(1)
            top = torch.topk(pr, 3)
            line += f"  | top rows {top.indices.tolist()} hold {float(top.values.sum() / pr.sum()):.1%} of the tensor's error"
            pc = d.pow(2).sum(0)
            topc = torch.topk(pc, 3)
            line += f", top cols {topc.indices.tolist()} {float(topc.values.sum() / pc.sum()):.1%}"
        print(line)
