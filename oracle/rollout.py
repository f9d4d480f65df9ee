"""Rollout loop restatements (TEST INFRASTRUCTURE, see oracle/__init__.py).

  * ``predict_ref``        - prediction/predict.py:16-42 (``predict``), kept for all clips
  * ``rollout_ref``        - batched sliding-window generalisation of the hot loop
                             prediction/predict.py:143-197
  * ``rollout_faithful``   - the literal B=1 loop of predict.py: first window is
                             [SOS, f1..f5] (:124-130, utils/sd_utils.py:31,151-153), then
                             all_latents = cat(inputs[:, :-1], preds) (:193), X = all[:, -5:] (:196)
  * ``chunked``            - B > 64 driver: the reference only accepts <= 64 clips per call
                             (PositionalEncoding max_len=64), so clip i sees PE[i mod 64].

``model`` is any callable with the reference contract
``model(src, tgt, tgt_mask) -> (S, B, E)`` plus ``get_tgt_mask(size)``.
"""
import torch

SOS_VALUE = 2.0          # utils/sd_utils.py:31
LATENT_SCALE = 0.18215   # utils/sd_utils.py:143,159


def predict_ref(model, window):
    """(B,W,E) -> (B,E): last-position prediction of every clip (predict.py:42 returns clip 0 only)."""
    with torch.no_grad():
        mask = model.get_tgt_mask(window.size(1)).to(window.device)
        pred = model(window, window, mask).permute(1, 0, 2)
        return pred[:, -1]


def predict_diff_ref(model, window):
    """prediction/predict_diff.py:14-41 for all clips: last-position output + second-to-last input frame (:33)."""
    return predict_ref(model, window) + window[:, -2]


def rollout_ref(model, ctx, n_pred, window, teacher=None, residual=False):
    """ctx (B,C,E) -> (B,n_pred,E).  ``teacher`` (B,n_pred,E): feed these instead of own predictions.
    ``residual``: the predict_diff.py variant of ``predict``."""
    allx, outs = ctx, []
    X = allx[:, -window:]
    for t in range(n_pred):
        p = predict_diff_ref(model, X) if residual else predict_ref(model, X)
        outs.append(p)
        nxt = p if teacher is None else teacher[:, t]
        allx = torch.cat([allx, nxt[:, None]], 1)
        X = allx[:, -window:]
    return torch.stack(outs, 1)


def rollout_faithful(model, frames, n_pred, residual=False):
    """frames (B,5,E) real latents. Reproduces prediction/predict.py:117-197 (any B<=64; reference B=1);
    ``residual`` = the same loop in prediction/predict_diff.py:117-192."""
    B, T, E = frames.shape
    sos = torch.full((B, 1, E), SOS_VALUE, dtype=frames.dtype)
    X = torch.cat([sos, frames], 1)              # use_sos=True, predict.py:124
    inputs = frames                               # predict.py:136-141 (all frames except SOS)
    preds = []
    for _ in range(n_pred):
        p = predict_diff_ref(model, X) if residual else predict_ref(model, X)   # predict.py:144 / predict_diff.py:140
        preds.append(p)
        all_latents = torch.cat([inputs[:, :-1], torch.stack(preds, 1)], 1)   # predict.py:193
        X = all_latents[:, -5:]                   # predict.py:196
    return torch.stack(preds, 1)


def chunked(fn, ctx, *args, chunk=64, **kw):
    """Apply a rollout function per <=64-clip chunk (the reference's batch limit) and concatenate."""
    outs = []
    for i in range(0, ctx.size(0), chunk):
        sub_kw = {k: (v[i:i + chunk] if torch.is_tensor(v) and v.size(0) == ctx.size(0) else v) for k, v in kw.items()}
        outs.append(fn(ctx[i:i + chunk], *args, **sub_kw))
    return torch.cat(outs, 0)


class FunctionalModel:
    """Adapter giving ``oracle.functional.forward`` the module call contract."""

    def __init__(self, sd, n_heads, operand="fp32", hp_first=False, dtype=torch.float32):
        from . import functional
        self._f = functional
        self.sd = {k: v.to(dtype) for k, v in sd.items()}
        self.n_heads, self.operand, self.hp_first, self.dtype = n_heads, operand, hp_first, dtype

    def get_tgt_mask(self, size):
        return self._f.causal_mask(size, self.dtype)

    def __call__(self, src, tgt, tgt_mask=None):
        return self._f.forward(self.sd, src.to(self.dtype), tgt.to(self.dtype), self.n_heads, tgt_mask,
                               operand=self.operand, hp_first=self.hp_first)


def max_rel_per_frame(ours, ref):
    """SURVEY.md 8d accuracy metric: max|ours-ref| / max|ref| per predicted frame. (B,P,E) -> (P,)"""
    num = (ours.double() - ref.double()).abs().amax(dim=(0, 2))
    den = ref.double().abs().amax(dim=(0, 2))
    return num / den
