"""Drop-in for the loss side of the reference's ``Trainer`` (trainers/trainer.py) - forward values, as used by its
validation loop (trainers/trainer.py:192-260) and for logging in the training loop (:168-178):

  * ``criterion(...)``                 - Trainer.criterion, trainers/trainer.py:88-109 (same keyword arguments)
  * ``gradient_difference_loss(x, y)`` - trainers/trainer.py:65-83
  * ``BiPatchNCE``                     - models/contrastive_loss.py:9-60 (same constructor; no (N*T, hw, hw) mask buffer)
  * ``validation_step(model, batch)``  - the body of validation_loop for one batch of latents (:203-224)

  * ``AdamTrainer``                    - opt = Adam(model.parameters(), lr) (:365) + the body of train_loop (:123-162):
                                         teacher-forced forward, criterion, backward, Adam step, and - one process per
                                         GPU - the NCCL all-reduce of the flat gradient vector in two buckets

All of them run hand-written CUDA through ``sdvg_criterion`` / ``sdvg_train_*`` (include/sdvg.h); CPU tensors raise."""
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


class _DeviceArray:
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def allreduce_buckets(flat, split, world_size, group=None, streams=None):
    """Sum-all-reduce ``flat[split:]`` then ``flat[:split]`` (the order the backward pass finishes them in).  Works on
    any backend (NCCL on the GPU; gloo on CPU tensors in the host-logic tests).  Returns the async work handles."""
    import torch.distributed as dist
    works = []
    for lo, hi in ((split, flat.numel()), (0, split)):
        if hi > lo:
            works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True))
    return works


class AdamTrainer:
    """``opt = optim.Adam(model.parameters(), lr=lr)`` (trainers/trainer.py:365) bound to the loss configuration of
    ``Trainer.criterion`` (:88) - ``step(new_batch)`` is one iteration of ``train_loop`` (:123-162).

    Data parallel (one process per GPU, ``torch.distributed`` initialised with NCCL): every rank passes its own shard
    of the global batch and ``pe_index`` = the clips' positions in the global batch; gradients are summed over ranks
    range by range as the backward pass finishes them - reduction and Adam update of a range run on a communication
    stream while the backward pass of the earlier layers continues - and Adam applies their mean on every rank, so all
    replicas stay identical.  (``overlap=False``: plain backward, two-bucket all-reduce, one Adam step.)  ``dropout`` defaults to the model's ``dropout_p``
    (see sdvg_train_set_dropout in include/sdvg.h); data-parallel ranks should pass different ``seed`` values."""

    def __init__(self, model, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, *, frames_to_predict=5, use_mse=True, use_L1=False,
                 use_gdl=True, lambda_gdl=1, alpha=2, use_contrastive=True, temperature=0.07, lambda_contrastive=0.1,
                 overlap=True, layers_per_bucket=3, dropout=None, seed=0):
        if use_mse and use_L1:
            raise RuntimeError("Invalid loss function combination")        # trainers/trainer.py:107-109
        # model.train() semantics: dropout with the model's DROPOUT_P at nn.Transformer's sites.  The masks come from the
        # library's counter-based generator (seed, step, site, element), not from torch's RNG stream.
        self.dropout = float(getattr(model, "dropout_p", 0.0) if dropout is None else dropout)
        self.seed = int(seed)
        if not 0.0 <= self.dropout < 1.0:
            raise RuntimeError("dropout probability must be in [0, 1)")
        self.model, self.lr, self.betas, self.eps = model, float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.loss_cfg = _lib.SdvgLossConfig(int(frames_to_predict), int(bool(use_mse)), int(bool(use_L1)), int(bool(use_gdl)),
                                            float(lambda_gdl), float(alpha), int(bool(use_contrastive)), float(temperature),
                                            float(lambda_contrastive))
        self.overlap = overlap
        self.layers_per_bucket = int(layers_per_bucket)
        self._comm_stream = None
        self._cb = None
        self._handle_seen = None
        self.steps = 0

    # -- views of library-owned memory
    def _handle(self, device):
        return self.model.engine(device)

    def gradients(self, device=None):
        """(flat fp32 gradient tensor [view], offset of the decoder-side bucket)."""
        device = device or next(self.model.parameters()).device
        h = self._handle(device)
        ptr, n, off = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(_lib.load().sdvg_train_gradients(h, C.byref(ptr), C.byref(n), C.byref(off)), h)
        return torch.as_tensor(_DeviceArray(ptr.value, n.value), device=device), int(off.value)

    def gradient(self, key, device=None):
        """Gradient of one state_dict entry, shaped like the parameter (a view)."""
        device = device or next(self.model.parameters()).device
        flat, _ = self.gradients(device)
        off, cnt = C.c_int64(), C.c_int64()
        _lib.check(_lib.load().sdvg_param_range(self._handle(device), key.encode(), C.byref(off), C.byref(cnt)))
        return flat[off.value: off.value + cnt.value].view(self.model.state_dict()[key].shape)

    def prediction(self, B, S_tgt, device=None):
        device = device or next(self.model.parameters()).device
        h = self._handle(device)
        ptr = C.c_void_p()
        _lib.check(_lib.load().sdvg_train_prediction(h, C.byref(ptr)), h)
        E = self.model.latent_dim
        return torch.as_tensor(_DeviceArray(ptr.value, S_tgt * B * E), device=device).view(S_tgt, B, E)

    def step(self, new_batch, pe_index=None):
        """One training iteration on latents ``new_batch`` (B, T+1, E) incl. the SOS frame (trainers/trainer.py:123-162).
        Returns the device tensor [total, mse, l1, gdl, contrastive] of this rank's shard."""
        import torch.distributed as dist
        if not new_batch.is_cuda:
            raise RuntimeError("sdvg_b200 training runs on CUDA only (no CPU fallback)")
        device = new_batch.device
        new_batch = new_batch.detach().float().contiguous()
        B, S, E = new_batch.shape
        y_input = new_batch[:, :-1].contiguous()                                   # :126
        y_expected = new_batch[:, 1:].permute(1, 0, 2).contiguous()                # :129-132
        self.model.reserve(max_clips=B, max_tokens=S)
        h = self._handle(device)
        if self.steps > 0 and self.model._handle_key != self._handle_seen:
            raise RuntimeError("the engine was rebuilt after training started (larger batch / window, new precision or device): "
                               "the Adam moments lived in the old engine; reserve() the largest shapes before the first step")
        self._handle_seen = self.model._handle_key       # (device, precision, limits, architecture) the engine was built for
        lib = _lib.load()
        _lib.check(lib.sdvg_train_set_dropout(h, self.dropout, self.seed), h)
        pe_ptr = None
        if pe_index is not None:
            pe_index = pe_index.to(device=device, dtype=torch.int32).contiguous()
            pe_ptr = pe_index.data_ptr()
        losses = torch.empty(5, device=device, dtype=torch.float32)
        stream = torch.cuda.current_stream(device)
        sp = C.c_void_p(stream.cuda_stream)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

        def backward(part):
            _lib.check(lib.sdvg_train_backward(h, new_batch.data_ptr(), y_input.data_ptr(), y_expected.data_ptr(), B, S, S - 1,
                                               C.byref(self.loss_cfg), pe_ptr, losses.data_ptr(), part, sp), h)

        adam = (self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / world)
        if self.overlap:
            # The library announces ranges of the flat gradient vector as soon as the backward pass - still being
            # enqueued - has finished them (sdvg_train_set_ready_callback, from the end of the vector to its start).  Each
            # range is all-reduced (N > 1) and Adam-updated on the communication stream while the backward pass of the
            # earlier layers keeps running on the caller's stream: neither touches the other's weights or gradients.
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device)
            comm = self._comm_stream
            flat = self.gradients(device)[0] if world > 1 else None
            state = {"first": 1, "error": None}

            def ready(_user, off, cnt):
                try:
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    comm.wait_event(ev)
                    with torch.cuda.stream(comm):
                        if world > 1:
                            dist.all_reduce(flat[off:off + cnt], op=dist.ReduceOp.SUM)      # stream-ordered on `comm`
                        _lib.check(lib.sdvg_train_adam_step_range(h, *adam, off, cnt, state["first"],
                                                                  C.c_void_p(comm.cuda_stream)), h)
                    state["first"] = 0
                except Exception as exc:            # an exception must not unwind through the C caller
                    state["error"] = state["error"] or exc

            self._cb = _lib.GRAD_READY_FN(ready)          # keep the ctypes thunk alive while the library holds it
            _lib.check(lib.sdvg_train_set_ready_callback(h, self._cb, None, self.layers_per_bucket), h)
            try:
                backward(0)
            finally:
                _lib.check(lib.sdvg_train_set_ready_callback(h, _lib.GRAD_READY_FN(0), None, 0), h)
            stream.wait_stream(comm)
            if state["error"] is not None:
                raise state["error"]
        else:
            backward(0)
            if world > 1:
                flat, split = self.gradients(device)
                for w in allreduce_buckets(flat, split, world):
                    w.wait()
            _lib.check(lib.sdvg_train_adam_step(h, *adam, sp), h)
        self.steps += 1
        self.model._pending_pull = self          # Transformer.state_dict() copies the trained weights back on demand
        return losses

    def pull_weights(self):
        """Copy the trained parameters back into the module (``model.state_dict()`` for torch.save, :294)."""
        self.model._pending_pull = self
        self.model._sync_trained_weights()
        return self.model
