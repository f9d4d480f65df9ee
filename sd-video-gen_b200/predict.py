"""Drop-in for the rollout interface of prediction/predict.py.

  * ``predict(model, input_sequence)``   - predict.py:16-42, same signature and return value (clip 0's
                                           last-position latent, shape (E,))
  * ``rollout(model, ctx, n_pred, ...)`` - the hot loop predict.py:143-197 for a batch of clips, on the device
  * ``rollout_from_host(...)``           - the same through host buffers (pinned H2D in, D2H out): the
                                           end-to-end call a user of the reference's script makes
"""
import torch

SOS_VALUE = 2.0          # utils/sd_utils.py:31
LATENT_SCALE = 0.18215   # utils/sd_utils.py:143,159


def predict(model, input_sequence):
    """predict.py:16-42.  input_sequence (B,S,E) on the model's CUDA device; returns pred[0,-1] (E,)."""
    model.eval()
    with torch.no_grad():
        pred = model(input_sequence, input_sequence, "causal")   # == get_tgt_mask(S).to(device), predict.py:24-26
        pred = pred.permute(1, 0, 2)
    return pred[0, -1]


def predict_diff(model, input_sequence):
    """prediction/predict_diff.py:14-41: residual prediction - the model output for the last position plus the
    second-to-last input frame (:33).  Returns pred[0,-1] like the reference."""
    model.eval()
    with torch.no_grad():
        pred = model(input_sequence, input_sequence, "causal").permute(1, 0, 2)
        last = pred[:, -1, :] + input_sequence[:, -2, :]
    return last[0]


def predict_future(model, input_sequence):
    """prediction/predict_future.py:16-42 (and the one-shot call at :156): no target mask, every position of the
    output is a predicted frame.  Returns pred[0,-1] like the reference; ``model(x, x, None)`` gives all frames."""
    model.eval()
    with torch.no_grad():
        pred = model(input_sequence, input_sequence, None).permute(1, 0, 2)
    return pred[0, -1]


def rollout(model, ctx, n_pred, window=5, *, use_sos=False, residual=False, teacher=None, pe_index=None,
            scale_in=1.0, scale_out=1.0, out=None):
    """ctx (B,C,E) device tensor -> (B,n_pred,E).  ``use_sos=True`` replays the literal predict.py sequence
    ([SOS,f1..f5] first, then the last 5 of [f1..f4,p1..pk]); otherwise a plain sliding window.
    ``residual=True`` is the predict_diff.py loop (each prediction += second-to-last window frame)."""
    model.eval()
    with torch.no_grad():
        return model.rollout(ctx, n_pred, window, faithful=use_sos, residual=residual, teacher=teacher,
                             pe_index=pe_index, scale_in=scale_in, scale_out=scale_out, out=out)


class HostRollout:
    """Reusable host<->device staging for ``rollout_from_host``: pinned input/output and device buffers are
    allocated once; each call does one H2D copy of the context, the device rollout, one D2H copy of the result."""

    def __init__(self, model, B, C, n_pred, device):
        E = model.latent_dim
        self.model, self.shape = model, (B, C, n_pred)
        self.ctx_pinned = torch.empty(B, C, E, dtype=torch.float32).pin_memory()
        self.out_pinned = torch.empty(B, n_pred, E, dtype=torch.float32).pin_memory()
        self.ctx_dev = torch.empty(B, C, E, dtype=torch.float32, device=device)
        self.out_dev = torch.empty(B, n_pred, E, dtype=torch.float32, device=device)
        self.h2d_bytes = self.ctx_pinned.numel() * 4
        self.d2h_bytes = self.out_pinned.numel() * 4

    def __call__(self, ctx_host=None, window=5, **kw):
        if ctx_host is not None:
            self.ctx_pinned.copy_(ctx_host)
        self.ctx_dev.copy_(self.ctx_pinned, non_blocking=True)
        rollout(self.model, self.ctx_dev, self.shape[2], window, out=self.out_dev, **kw)
        self.out_pinned.copy_(self.out_dev, non_blocking=True)
        torch.cuda.current_stream(self.ctx_dev.device).synchronize()
        return self.out_pinned


def rollout_from_host(model, ctx_host, n_pred, window=5, device=None, **kw):
    device = device or next(model.parameters()).device
    B, C, _ = ctx_host.shape
    return HostRollout(model, B, C, n_pred, device)(ctx_host, window, **kw).clone()
