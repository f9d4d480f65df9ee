"""Restatement of one iteration of the reference's training loop (TEST INFRASTRUCTURE, see oracle/__init__.py).

``train_step_ref`` follows Trainer.train_loop, trainers/trainer.py:123-162:
    y_input = new_batch[:, :-1]; y_expected = new_batch[:, 1:].permute(1, 0, 2)         (:126-132)
    pred = model(new_batch, y_input, model.get_tgt_mask(S_tgt))                           (:136-141)
    loss = loss_fn(pred[-P:], y_expected[-P:])                                            (:145)
    opt.zero_grad(); loss.backward(); opt.step()                                          (:160-162)
with ``opt = torch.optim.Adam(model.parameters(), lr)`` (:365) and ``loss_fn`` = oracle.losses.criterion.
The model must be built with dropout_p = 0 (the reference trains with DROPOUT_P 0.1, whose masks come from
torch's RNG stream and are not reproducible by another implementation).
Data-parallel oracle: the same step on the concatenated global batch (equal shards, mean-reduced losses).
Pinned by tests/golden/train_step.npz (oracle/make_golden_train.py: the reference's own Trainer.criterion,
BiPatchNCE and Transformer under autograd)."""
import torch

from . import losses


def train_step_ref(model, opt, new_batch, frames_to_predict, **loss_kw):
    """Returns (loss, pred.detach(), {param name: grad clone}); updates `model` in place through `opt`."""
    model.train()
    y_input = new_batch[:, :-1]
    y_expected = new_batch[:, 1:].permute(1, 0, 2)
    mask = model.get_tgt_mask(y_input.size(1))
    pred = model(new_batch, y_input, mask)
    loss = losses.criterion(**loss_kw)(pred[-frames_to_predict:], y_expected[-frames_to_predict:])
    opt.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    return loss.detach(), pred.detach(), grads


def make_batch(B, S, E, seed, sos=True):
    """Synthetic latents shaped like encode_batch(..., use_sos=True) output (utils/sd_utils.py:151-153): an SOS
    frame of 2.0 followed by S-1 unit-variance frames."""
    x = torch.randn(B, S, E, generator=torch.Generator().manual_seed(seed))
    if sos:
        x[:, 0] = 2.0
    return x


def train_grads_functional(state_dict, n_heads, new_batch, frames_to_predict, pe_index=None, drop=None, return_dpred=False,
                           **loss_kw):
    """Same iteration on the op-by-op restatement (oracle/functional.py) under autograd - used where hooks are needed:
    ``pe_index`` (data-parallel shards) and ``drop`` (an oracle.dropout.Dropper: training-mode dropout with the
    library's own masks, DROPOUT_P of the reference's configs).  Returns (loss, pred, {name: grad}) for the float
    parameters of ``state_dict`` (the positional table is a buffer and gets no gradient); with ``return_dpred`` also
    dL/dpred (S_tgt, B, E) - the upstream gradient ``loss.backward()`` hands to the model's backward pass."""
    from . import functional as F
    sd = {k: v.detach().clone() for k, v in state_dict.items()}
    names = [k for k in sd if k != "positional_encoder.pos_encoding"]
    for k in names:
        sd[k].requires_grad_(True)
    y_input = new_batch[:, :-1]
    y_expected = new_batch[:, 1:].permute(1, 0, 2)
    pred = F.forward(sd, new_batch, y_input, n_heads, F.causal_mask(y_input.size(1), new_batch.dtype), pe_index=pe_index, drop=drop)
    if return_dpred:
        pred.retain_grad()
    loss = losses.criterion(**loss_kw)(pred[-frames_to_predict:], y_expected[-frames_to_predict:])
    loss.backward()
    grads = {k: sd[k].grad.detach().clone() for k in names}
    if return_dpred:
        return loss.detach(), pred.detach(), grads, pred.grad.detach().clone()
    return loss.detach(), pred.detach(), grads
