"""Restatement of the reference's loss functions (TEST INFRASTRUCTURE, see oracle/__init__.py).

  * ``gdl``        - Trainer.gradient_difference_loss, trainers/trainer.py:65-83
  * ``bipatch_nce``- BiPatchNCE.forward, models/contrastive_loss.py:28-60, including the diag / non-diag split at
                     :39-47 (the scores are gt_f . pred_f^T / temperature either way; the split stops the gradient
                     of the direction-1 negatives, which matters for oracle/train.py)
  * ``criterion``  - Trainer.criterion, trainers/trainer.py:88-109
Inputs are (P, B, E) sequence-first slices ``pred[-P:]`` / ``y_expected[-P:]`` like the reference's call sites
(trainers/trainer.py:145, :224).  Pinned by tests/golden/losses.npz (oracle/make_golden_losses.py)."""
import math

import torch


def gdl(x, y, alpha=1.0):
    P, B, E = x.shape
    s = int(math.isqrt(E // 4))
    X = x.reshape(P, B, 4, s, s)
    Y = y.reshape(P, B, 4, s, s)
    v = ((X[..., 1:, :] - X[..., :-1, :]).abs() - (Y[..., 1:, :] - Y[..., :-1, :]).abs()).abs()
    h = ((X[..., :, 1:] - X[..., :, :-1]).abs() - (Y[..., :, 1:] - Y[..., :, :-1]).abs()).abs()
    return (v.pow(alpha).sum() + h.pow(alpha).sum()) / x.numel()


def bipatch_nce(x, y, temperature=0.07):
    """x = prediction, y = ground truth, both (P, B, E)."""
    P, B, E = x.shape
    hw = E // 4
    pred_f = x.permute(1, 0, 2).reshape(B * P, 4, hw).transpose(1, 2)   # (N T) (h w) C
    gt_f = y.permute(1, 0, 2).reshape(B * P, 4, hw).transpose(1, 2)
    eye = torch.eye(hw, dtype=x.dtype, device=x.device)
    s1 = (gt_f @ pred_f.transpose(1, 2) * eye + gt_f @ pred_f.detach().transpose(1, 2) * (1.0 - eye)) / temperature   # :39-42
    s2 = (pred_f @ gt_f.transpose(1, 2) * eye + pred_f @ gt_f.detach().transpose(1, 2) * (1.0 - eye)) / temperature   # :45-48
    diag = torch.arange(hw)
    l1 = (torch.logsumexp(s1, -1) - s1[:, diag, diag]).mean()            # CrossEntropyLoss vs identity target, :56
    l2 = (torch.logsumexp(s2, -1) - s2[:, diag, diag]).mean()            # :57
    return 0.5 * (l1 + l2)


def criterion(use_mse=True, use_L1=False, use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07,
              lambda_contrastive=0.1):
    if use_mse and use_L1:
        return None                                                      # trainers/trainer.py:107-109
    def loss(x, y):
        t = x.new_zeros(())
        if use_mse:
            t = t + ((x - y) ** 2).mean()
        if use_L1:
            t = t + (x - y).abs().mean()
        if use_gdl:
            t = t + lambda_gdl * gdl(x, y, alpha)
        if use_contrastive:
            t = t + lambda_contrastive * bipatch_nce(x, y, temperature)
        return t
    return loss
