"""Golden values of the reference's loss functions (trainers/trainer.py:65-109, models/contrastive_loss.py) as used
by its validation / training loops.  TEST INFRASTRUCTURE (see oracle/__init__.py); run in the build container:

    python oracle/make_golden_losses.py        # writes tests/golden/losses.npz

trainers/trainer.py imports utils.sd_utils -> diffusers at module top (absent here), so diffusers / wandb are
stubbed with MagicMock before importing - SURVEY.md 8c; the loss code itself is the reference's, unmodified."""
import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")


def main():
    for mod in ("diffusers", "diffusers.schedulers", "diffusers.schedulers.scheduling_ddim", "wandb", "cv2"):
        if mod not in sys.modules:
            sys.modules[mod] = MagicMock()
    sys.argv = ["x", "--dataset", "ball", "--config", "1_17_ball_complex_L1_64"]
    sys.path.insert(0, REF)
    os.chdir(REF)
    from trainers.trainer import Trainer          # the reference, unmodified
    from models.contrastive_loss import BiPatchNCE

    out = {}
    g = torch.Generator().manual_seed(77)
    for tag, (P, B, F) in {"f64": (5, 3, 64), "f128": (5, 2, 128), "p1": (1, 4, 64)}.items():
        h = F // 8
        E = 4 * h * h
        x = torch.randn(P, B, E, generator=g)
        y = x + 0.3 * torch.randn(P, B, E, generator=g)
        tr = Trainer.__new__(Trainer)
        tr.config = SimpleNamespace(FRAMES_TO_PREDICT=[P], BATCH_SIZE=[B], FRAME_SIZE=F)
        tr.device = torch.device("cpu")
        out[f"{tag}.x"], out[f"{tag}.y"] = x, y
        out[f"{tag}.shape"] = np.array([P, B, F])
        out[f"{tag}.mse"] = torch.nn.MSELoss()(x, y)
        out[f"{tag}.l1"] = torch.nn.L1Loss()(x, y)
        out[f"{tag}.gdl1"] = tr.gradient_difference_loss(x, y, 1)
        out[f"{tag}.gdl2"] = tr.gradient_difference_loss(x, y, 2)
        nce = BiPatchNCE(N=B, T=P, h=h, w=h, temperature=0.07)
        out[f"{tag}.nce"] = nce(x.permute(1, 0, 2).reshape(-1, P, 4, h, h), y.permute(1, 0, 2).reshape(-1, P, 4, h, h))
        # criterion combinations used by the BASELINE configs: L1 only (C1-C4), MSE+GDL(alpha 2)+0.1*NCE (C5)
        out[f"{tag}.crit_l1"] = tr.criterion(use_mse=False, use_L1=True, use_gdl=False, use_contrastive=False)(x, y)
        out[f"{tag}.crit_c5"] = tr.criterion(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2,
                                             use_contrastive=True, temperature=0.07, lambda_contrastive=0.1)(x, y)
        out[f"{tag}.crit_all"] = tr.criterion(use_mse=False, use_L1=True, use_gdl=True, lambda_gdl=0.5, alpha=1,
                                              use_contrastive=True, temperature=0.1, lambda_contrastive=0.0258)(x, y)
    path = os.path.join(OUT, "losses.npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path) / 1e3, "KB")
    for k, v in out.items():
        if not k.endswith((".x", ".y", ".shape")):
            print(k, float(v))


if __name__ == "__main__":
    main()
