"""Drop-in for the loss side of the reference's ``Trainer`` (trainers/trainer.py) - forward values, as used by its
validation loop (trainers/trainer.py:192-260) and for logging in the training loop (:168-178):

  * ``criterion(...)``                 - Trainer.criterion, trainers/trainer.py:88-109 (same keyword arguments)
  * ``gradient_difference_loss(x, y)`` - trainers/trainer.py:65-83
  * ``BiPatchNCE``                     - models/contrastive_loss.py:9-60 (same constructor; no (N*T, hw, hw) mask buffer)
  * ``validation_step(model, batch)``  - the body of validation_loop for one batch of latents (:203-224)

All of them run hand-written CUDA through ``sdvg_criterion`` (include/sdvg.h); CPU tensors raise.  The backward
pass / optimiser step of the training loop are not part of this build."""
import ctypes as C
import math

import torch

from . import _lib


def _evaluate(x, y, *, use_mse, use_l1, use_gdl, lambda_gdl, alpha, use_contrastive, temperature, lambda_contrastive):
    if not x.is_cuda or not y.is_cuda:
        raise RuntimeError("sdvg_b200 losses run on CUDA only (no CPU fallback)")
    if x.shape != y.shape or x.dim() != 3:
        raise RuntimeError(f"expected two (P, B, E) tensors of equal shape, got {tuple(x.shape)} and {tuple(y.shape)}")
    P, B, E = x.shape
    s = int(math.isqrt(E // 4))
    if 4 * s * s != E:
        raise RuntimeError(f"E = {E} is not 4 * h * h")
    x = x.detach().float().contiguous()
    y = y.detach().float().contiguous()
    out = torch.empty(5, device=x.device, dtype=torch.float32)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    rc = _lib.load().sdvg_criterion(x.device.index or 0, x.data_ptr(), y.data_ptr(), P, B, s, s, int(bool(use_mse)),
                                    int(bool(use_l1)), int(bool(use_gdl)), float(lambda_gdl), float(alpha),
                                    int(bool(use_contrastive)), float(temperature), float(lambda_contrastive),
                                    out.data_ptr(), C.c_void_p(stream))
    _lib.check(rc)
    return out          # total, mse, l1, gdl, contrastive


def criterion(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07,
              lambda_contrastive=0.1):
    """trainers/trainer.py:88-109: returns ``loss(x, y)`` for (P,B,E) slices, or None for the invalid MSE+L1 combination."""
    if use_mse and use_L1:
        print("Invalid loss function combination")
        return None
    def loss(x, y):
        return _evaluate(x, y, use_mse=use_mse, use_l1=use_L1, use_gdl=use_gdl, lambda_gdl=lambda_gdl, alpha=alpha,
                         use_contrastive=use_contrastive, temperature=temperature,
                         lambda_contrastive=lambda_contrastive)[0]
    return loss


def loss_terms(x, y, alpha=2, temperature=0.07):
    """All individual terms in one pass: dict(mse, l1, gdl, contrastive) - what validation_loop logs (:236-246)."""
    o = _evaluate(x, y, use_mse=True, use_l1=False, use_gdl=True, lambda_gdl=1.0, alpha=alpha, use_contrastive=True,
                  temperature=temperature, lambda_contrastive=1.0)
    return {"mse": o[1], "l1": o[2], "gdl": o[3], "contrastive": o[4]}


def gradient_difference_loss(frameX_flattened, frameY_flattened, alpha=1):
    """trainers/trainer.py:65-83."""
    return _evaluate(frameX_flattened, frameY_flattened, use_mse=False, use_l1=False, use_gdl=True, lambda_gdl=1.0,
                     alpha=alpha, use_contrastive=False, temperature=1.0, lambda_contrastive=0.0)[3]


class BiPatchNCE(torch.nn.Module):
    """models/contrastive_loss.py:9-60.  forward(pred_f, gt_f) with (N, T, C=4, h, w) tensors, like the reference."""

    def __init__(self, N, T, h, w, temperature=0.07):
        super().__init__()
        self.N, self.T, self.h, self.w, self.temperature = N, T, h, w, temperature

    def forward(self, pred_f, gt_f):
        N, T, Cc, h, w = pred_f.shape
        if Cc != 4:
            raise RuntimeError("BiPatchNCE kernel expects 4 latent channels")
        # (N, T, C, h, w) -> (T, N, C*h*w): the layout the criterion kernel reads (trainers/trainer.py:106 in reverse)
        x = pred_f.reshape(N, T, -1).permute(1, 0, 2)
        y = gt_f.reshape(N, T, -1).permute(1, 0, 2)
        return _evaluate(x, y, use_mse=False, use_l1=False, use_gdl=False, lambda_gdl=0.0, alpha=1.0,
                         use_contrastive=True, temperature=self.temperature, lambda_contrastive=1.0)[4]


def validation_step(model, new_batch, frames_to_predict, loss_fn):
    """One iteration of validation_loop (trainers/trainer.py:203-224) on a batch of latents (B, T+1, E) incl. SOS:
    teacher-forced forward with S_src = T+1, S_tgt = T and the causal mask, loss on the last `frames_to_predict`
    positions.  Returns (loss, pred) with pred (T, B, E)."""
    model.eval()
    with torch.no_grad():
        y_input = new_batch[:, :-1]
        y_expected = new_batch[:, 1:].permute(1, 0, 2)
        pred = model(new_batch, y_input.contiguous(), "causal")
        loss = loss_fn(pred[-frames_to_predict:], y_expected[-frames_to_predict:].contiguous())
    return loss, pred
