"""Golden values of one training iteration of the reference (trainers/trainer.py:123-162) - its own Transformer
(models/transformer.py, dropout_p = 0), its own Trainer.criterion / gradient_difference_loss / BiPatchNCE under
torch autograd, and torch.optim.Adam as constructed at trainers/trainer.py:365.  TEST INFRASTRUCTURE; run in the
build container:

    python oracle/make_golden_train.py        # writes tests/golden/train_step.npz

The model is small (d128 H4 1e/2d, E256, weights by seed + checksum); per parameter the fixture keeps the gradient's
max |.|, sum, and a strided sample, plus samples of the weights after two Adam steps and the loss values."""
import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, os.path.join(HERE, ".."))

ARCH = dict(d=128, H=4, Le=1, Ld=2)
SEED = 11
B, S, P = 3, 6, 5
LR = 1e-3     # large enough for the parameter change of two steps to be far above fp32 rounding
CASES = {
    "c5": dict(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07,
               lambda_contrastive=0.1),                                   # config 11_19_wallpushups_all_losses_test
    "l1": dict(use_mse=False, use_L1=True, use_gdl=False, use_contrastive=False),     # configs C1-C4 (USE_L1 only)
    "gdl1": dict(use_mse=False, use_L1=True, use_gdl=True, lambda_gdl=0.5, alpha=1, use_contrastive=True, temperature=0.1,
                 lambda_contrastive=0.05),
}


def sample(t, n=64):
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx]


def main():
    for mod in ("diffusers", "diffusers.schedulers", "diffusers.schedulers.scheduling_ddim", "wandb", "cv2"):
        if mod not in sys.modules:
            sys.modules[mod] = MagicMock()
    sys.argv = ["x", "--dataset", "ball", "--config", "1_17_ball_complex_L1_64"]
    sys.path.insert(0, REF)
    os.chdir(REF)
    from trainers.trainer import Trainer          # the reference, unmodified
    from models.transformer import Transformer
    from oracle.make_golden import sd_checksum
    from oracle.train import make_batch

    torch.set_num_threads(8)
    out = {"arch": np.array([ARCH["d"], ARCH["H"], ARCH["Le"], ARCH["Ld"], 256]), "seed": SEED, "shape": np.array([B, S, P]),
           "lr": LR}
    for tag, kw in CASES.items():
        torch.manual_seed(SEED)
        model = Transformer(0, ARCH["d"], ARCH["H"], ARCH["Le"], ARCH["Ld"], 0.0)     # dropout_p = 0
        out["checksum"] = np.array(sd_checksum(model.state_dict()))
        tr = Trainer.__new__(Trainer)
        tr.config = SimpleNamespace(FRAMES_TO_PREDICT=[P], BATCH_SIZE=[B], FRAME_SIZE=64)
        tr.device = torch.device("cpu")
        loss_fn = tr.criterion(**kw)
        opt = torch.optim.Adam(model.parameters(), lr=LR)                               # trainers/trainer.py:365
        model.train()
        for step in range(2):
            new_batch = make_batch(B, S, 256, seed=100 + step)
            y_input = new_batch[:, :-1]                                                 # :126
            y_expected = new_batch[:, 1:].permute(1, 0, 2)                              # :129-132
            tgt_mask = model.get_tgt_mask(y_input.size(1))                              # :135-136
            pred = model(new_batch, y_input, tgt_mask)                                  # :141
            loss = loss_fn(pred[-P:], y_expected[-P:])                                  # :145
            opt.zero_grad(); loss.backward(); opt.step()                                # :160-162
            out[f"{tag}.loss{step}"] = loss.detach()
            if step == 0:
                out[f"{tag}.pred0"] = pred.detach()
                for k, p in model.named_parameters():
                    out[f"{tag}.g.{k}.amax"] = p.grad.abs().max()
                    out[f"{tag}.g.{k}.sum"] = p.grad.double().sum()
                    out[f"{tag}.g.{k}.s"] = sample(p.grad)
        for k, p in model.named_parameters():
            out[f"{tag}.w.{k}.s"] = sample(p)
        print(tag, float(out[f"{tag}.loss0"]), float(out[f"{tag}.loss1"]))
    path = os.path.join(OUT, "train_step.npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path) / 1e3, "KB")


if __name__ == "__main__":
    main()
