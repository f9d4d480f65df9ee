"""Drop-in for the reference's ``models.transformer.Transformer`` (models/transformer.py:9-93).

Same positional constructor, same sub-module names (``positional_encoder``, ``embedding``, ``transformer``,
``out``) and therefore the same ``state_dict`` keys and shapes, same ``forward(src, tgt, tgt_mask)`` contract
(returns ``(S_tgt, B, E)``), same ``get_tgt_mask`` / ``create_pad_mask``.  The arithmetic is not PyTorch's:
``forward`` hands raw device pointers to libsdvg.so (hand-written sm_100a kernels, include/sdvg.h).  The
``nn.Transformer`` / ``nn.Linear`` children are parameter containers only (initialisation, ``state_dict``,
``load_state_dict``, ``.to()``); they are never called.

There is no CPU path: CPU tensors, a missing library or a non-sm_100 device raise ``RuntimeError``.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib
from .config import latent_dim
from .positional_encoding import PositionalEncoding


class _TrainForward(torch.autograd.Function):
    """``pred = model(src, tgt, mask)`` in train() mode as a node of torch's autograd graph: forward = libsdvg's
    training forward pass (activations saved in the engine), backward = libsdvg's backward pass from dL/dpred; the
    parameter gradients come back as tensors, so ``loss.backward()`` fills ``p.grad`` and ``optim.Adam(model.parameters())``
    steps exactly as in the reference's loop (trainers/trainer.py:141-165,365)."""

    @staticmethod
    def forward(ctx, model, src, tgt, pe_index, *params):
        lib = _lib.load()
        device = src.device
        B, Ss, St = src.size(0), src.size(1), tgt.size(1)
        h = model.engine(device)
        _lib.check(lib.sdvg_train_set_dropout(h, float(model.dropout_p), int(getattr(model, "dropout_seed", 0))), h)
        pred = torch.empty(St, B, model.latent_dim, device=device, dtype=torch.float32)
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(lib.sdvg_train_forward(h, src.data_ptr(), tgt.data_ptr(), B, Ss, St,
                                          None if pe_index is None else pe_index.data_ptr(), pred.data_ptr(), stream), h)
        ctx.model, ctx.handle, ctx.keep = model, h, (src, tgt, pe_index)     # src / tgt are read again by the backward pass
        ctx.keys = [k for k, _ in model.named_parameters() if k != "learned_tgt"]
        return pred

    @staticmethod
    def backward(ctx, dpred):
        lib = _lib.load()
        model, h = ctx.model, ctx.handle
        if model._handle is not h:
            raise RuntimeError("the engine was rebuilt between forward and backward (reserve / set_precision / .to())")
        device = dpred.device
        dpred = dpred.detach().float().contiguous()
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(lib.sdvg_train_backward_from(h, dpred.data_ptr(), stream), h)
        ptr, n, off = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(lib.sdvg_train_gradients(h, C.byref(ptr), C.byref(n), C.byref(off)), h)

        class _Arr:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<f4", "data": (int(ptr.value), False),
                                        "version": 3, "strides": None}
        flat = torch.as_tensor(_Arr(), device=device)
        grads = []
        sd = dict(model.named_parameters())
        for k in ctx.keys:
            o, c = C.c_int64(), C.c_int64()
            _lib.check(lib.sdvg_param_range(h, k.encode(), C.byref(o), C.byref(c)))
            # a copy: the flat vector is library memory that the next backward pass overwrites
            grads.append(flat[o.value:o.value + c.value].view(sd[k].shape).clone())
        return (None, None, None, None, *grads)


class Transformer(nn.Module):
    def __init__(self, num_tokens=0, dim_model=256, num_heads=8, num_encoder_layers=6, num_decoder_layers=6,
                 dropout_p=0.1, *, frame_size=None, precision="fp32", max_clips=64, max_tokens=16, max_history=32):
        super().__init__()
        if frame_size is None:
            # reference behaviour (models/transformer.py:23,28-29): FRAME_SIZE comes from argv + ./config/<name>.yml
            from .config import parse_config_args
            self.config, self.args = parse_config_args()
            frame_size = self.config.FRAME_SIZE
        else:
            self.config, self.args = None, None
        self.dim_model = dim_model
        self.num_heads = num_heads
        self.height = frame_size
        self.width = frame_size
        self.compression = 8
        self.latent_dim = latent_dim(frame_size, self.compression)
        # creation order = the reference's, so torch.manual_seed(s) gives identical random-init weights
        self.positional_encoder = PositionalEncoding(dim_model=dim_model, dropout_p=dropout_p, max_len=64)
        self.embedding = nn.Linear(self.latent_dim, dim_model)
        self.transformer = nn.Transformer(d_model=dim_model, nhead=num_heads, num_encoder_layers=num_encoder_layers,
                                          num_decoder_layers=num_decoder_layers, dropout=dropout_p)
        self.out = nn.Linear(dim_model, self.latent_dim)
        self.dropout_p = dropout_p
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self.precision = precision
        self._limits = dict(max_clips=max_clips, max_tokens=max_tokens, max_history=max_history)
        self._handle = None
        self._handle_key = None
        self._weights_stamp = None

    # ------------------------------------------------------------------ engine plumbing
    def _arch(self):
        t = self.transformer
        return (self.dim_model, self.num_heads, len(t.encoder.layers), len(t.decoder.layers), self.latent_dim,
                t.encoder.layers[0].linear1.out_features if len(t.encoder.layers) else
                t.decoder.layers[0].linear1.out_features)

    def state_dict(self, *args, **kwargs):
        """nn.Module.state_dict; after AdamTrainer steps the trained values live in the engine, so they are copied back
        into the parameters first (torch.save(model.state_dict()), trainers/trainer.py:294, sees the trained weights)."""
        self._sync_trained_weights()
        return super().state_dict(*args, **kwargs)

    def _sync_trained_weights(self):
        """If an AdamTrainer (or loss.backward() + optimizer bridge) left newer weights in the engine than in the
        module's parameters, copy them back NOW - while the handle that holds them is still alive.  Called before the
        engine is freed or rebuilt (reserve(), set_precision(), .to()), before any weight push and by state_dict()."""
        if getattr(self, "_pending_pull", None) is None:
            return
        self._pending_pull = None
        if self._handle is not None:
            self._pull_from_handle(self._handle)

    def _pull_from_handle(self, handle):
        lib = _lib.load()
        same_place = True
        dev_index = self._handle_key[0] if self._handle_key else torch.cuda.current_device()
        stream = C.c_void_p(torch.cuda.current_stream(dev_index).cuda_stream)   # ordered after the training kernels
        with torch.no_grad():
            for k, p in self.named_parameters():
                if k == "learned_tgt":
                    continue
                # the library copies with cudaMemcpyDefault: the destination may be on any device or on the host
                buf = torch.empty(p.shape, dtype=torch.float32, device=p.device)
                _lib.check(lib.sdvg_get_weight(handle, k.encode(), C.c_void_p(buf.data_ptr()), stream), handle)
                p.copy_(buf)
                same_place = same_place and p.is_cuda
        if same_place and self._handle is handle:
            self._weights_stamp = self._stamp()     # the engine already holds exactly these values

    def load_state_dict(self, *args, **kwargs):
        self._pending_pull = None            # loaded values supersede whatever an AdamTrainer left in the engine
        return super().load_state_dict(*args, **kwargs)

    def _stamp(self):
        sd = super().state_dict()       # not self.state_dict(): that one pulls trained weights back from the engine
        return tuple((k, v.data_ptr(), v._version) for k, v in sd.items())

    def _free(self):
        if self._handle is not None:
            self._sync_trained_weights()          # trained weights live only in the engine: rescue them first
            _lib.load().sdvg_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def set_precision(self, precision):
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        if precision != self.precision:
            self.precision = precision
            self._free()

    def reserve(self, max_clips=None, max_tokens=None, max_history=None):
        """Grow the workspace limits (workspace is allocated once, outside the hot path)."""
        for k, v in (("max_clips", max_clips), ("max_tokens", max_tokens), ("max_history", max_history)):
            if v is not None and v > self._limits[k]:
                self._limits[k] = int(v)
                self._free()

    def engine(self, device):
        """The libsdvg handle for `device`, (re)created and weight-synced on demand."""
        if device.type != "cuda":
            raise RuntimeError("sdvg_b200.Transformer runs on CUDA only (no CPU fallback); move the module and "
                               "its inputs to a B200 with .to('cuda')")
        lib = _lib.load()
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        key = (dev_index, self.precision, tuple(sorted(self._limits.items())), self._arch())
        if self._handle is None or self._handle_key != key:
            self._free()                          # (rescues trained weights out of the old engine first)
            d, H, Le, Ld, E, ff = self._arch()
            cfg = _lib.SdvgConfig(d, H, Le, Ld, E, ff, 1e-5, self._limits["max_clips"], self._limits["max_tokens"],
                                  self._limits["max_history"], _lib.PRECISIONS[self.precision], dev_index)
            h = C.c_void_p()
            _lib.check(lib.sdvg_create(C.byref(cfg), C.byref(h)))
            self._handle, self._handle_key, self._weights_stamp = h, key, None
        stamp = self._stamp()
        if stamp != self._weights_stamp and getattr(self, "_pending_pull", None) is not None:
            # the parameters moved (.to(), in-place edit) while newer weights still sit in the engine: never push stale
            # values over what an optimizer step left there - bring the trained weights home first
            self._sync_trained_weights()
            stamp = self._stamp()
        if stamp != self._weights_stamp:
            for k, v in super().state_dict().items():
                if k == "learned_tgt":      # TransformerFuture's extra parameter is not used by forward
                    continue
                t = v.detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.float().contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(lib.sdvg_set_weight(self._handle, k.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()),
                           self._handle)
            stream = torch.cuda.current_stream(dev_index).cuda_stream
            _lib.check(lib.sdvg_finalize_weights(self._handle, C.c_void_p(stream)), self._handle)
            self._weights_stamp = stamp
        return self._handle

    @staticmethod
    def _f32c(t, device):
        if t.device != device:
            raise RuntimeError(f"tensor on {t.device}, model inputs must be on {device}")
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.float().contiguous()
        return t

    def _check_eval(self):
        if self.training and self.dropout_p > 0:
            raise RuntimeError("rollout()/predict() are the inference path: call model.eval() first (as prediction/predict.py:17 "
                               "does).  Training goes through model(...) + loss.backward() in train() mode, or the fused "
                               "sdvg_b200.AdamTrainer.step()")

    # ------------------------------------------------------------------ reference API
    def forward(self, src, tgt, tgt_mask=None, src_pad_mask=None, tgt_pad_mask=None, *, pe_index=None):
        """models/transformer.py:47-68.  src (B,S_src,E), tgt (B,S_tgt,E) -> (S_tgt,B,E).

        tgt_mask: None, the string "causal" (== get_tgt_mask(S_tgt) without building it), or an additive float
        (S_tgt,S_tgt) tensor as the reference passes.  pe_index (B,) int32 overrides the batch-position PE row
        (extension; needed for B > 64, where the reference itself raises)."""
        if src_pad_mask is not None or tgt_pad_mask is not None:
            raise RuntimeError("padding masks are None at every reference call site and are not supported")
        if self.training:
            return self._forward_train(src, tgt, tgt_mask, pe_index)
        if src.dim() != 3 or tgt.dim() != 3 or src.size(0) != tgt.size(0) or src.size(2) != self.latent_dim \
                or tgt.size(2) != self.latent_dim:
            raise RuntimeError(f"expected src (B,S,{self.latent_dim}) and tgt (B,S',{self.latent_dim}), got "
                               f"{tuple(src.shape)} and {tuple(tgt.shape)}")
        device = src.device
        B, Ss, St = src.size(0), src.size(1), tgt.size(1)
        if pe_index is None and B > 64:
            # the reference: "The size of tensor a (B) must match the size of tensor b (64)" (positional_encoding.py:35)
            raise RuntimeError(f"The size of tensor a ({B}) must match the size of tensor b (64) at non-singleton "
                               "dimension 0 (positional encoding is indexed by batch position; pass pe_index for B > 64)")
        self.reserve(max_clips=B, max_tokens=max(Ss, St))
        h = self.engine(device)
        same = src is tgt or (src.data_ptr() == tgt.data_ptr() and src.shape == tgt.shape and src.stride() == tgt.stride())
        s = self._f32c(src, device)
        t = s if same else self._f32c(tgt, device)
        mask_kind, mask_ptr, keep = 0, None, None
        if isinstance(tgt_mask, str):
            if tgt_mask != "causal":
                raise ValueError("tgt_mask string must be 'causal'")
            mask_kind = 1
        elif tgt_mask is not None:
            if tuple(tgt_mask.shape) != (St, St):
                raise RuntimeError(f"tgt_mask must be ({St},{St})")
            keep = tgt_mask.to(device=device, dtype=torch.float32).contiguous()
            mask_kind, mask_ptr = 2, keep.data_ptr()
        pe_ptr = None
        if pe_index is not None:
            pe_index = pe_index.to(device=device, dtype=torch.int32).contiguous()
            if pe_index.numel() != B:
                raise RuntimeError("pe_index must have one entry per clip")
            pe_ptr = pe_index.data_ptr()
        out = torch.empty(St, B, self.latent_dim, device=device, dtype=torch.float32)
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.load().sdvg_forward(h, s.data_ptr(), t.data_ptr(), B, Ss, St, mask_kind, mask_ptr, pe_ptr,
                                            out.data_ptr(), C.c_void_p(stream)), h)
        return out

    def _forward_train(self, src, tgt, tgt_mask, pe_index):
        """model.train(): the call of trainers/trainer.py:141 - dropout active (the library's counter-based masks,
        ``self.dropout_seed``), activations saved, and the result carries a grad_fn so that the reference's own
        ``loss.backward()`` / ``optim.Adam(model.parameters()).step()`` work unchanged.  The fused fast path for the
        same iteration is ``sdvg_b200.AdamTrainer.step``."""
        if src.dim() != 3 or tgt.dim() != 3 or src.size(0) != tgt.size(0) or src.size(2) != self.latent_dim \
                or tgt.size(2) != self.latent_dim:
            raise RuntimeError(f"expected src (B,S,{self.latent_dim}) and tgt (B,S',{self.latent_dim}), got "
                               f"{tuple(src.shape)} and {tuple(tgt.shape)}")
        device = src.device
        B, Ss, St = src.size(0), src.size(1), tgt.size(1)
        if pe_index is None and B > 64:
            raise RuntimeError(f"The size of tensor a ({B}) must match the size of tensor b (64) at non-singleton "
                               "dimension 0 (positional encoding is indexed by batch position; pass pe_index for B > 64)")
        causal = isinstance(tgt_mask, str) and tgt_mask == "causal"
        if not causal:
            # every training call site of the reference passes get_tgt_mask(T) (trainers/trainer.py:137-141)
            if tgt_mask is None or tuple(tgt_mask.shape) != (St, St) or not torch.equal(tgt_mask.detach().cpu().float(),
                                                                                        self.get_tgt_mask(St)):
                raise RuntimeError("the training forward pass implements the causal target mask of get_tgt_mask() "
                                   "(the only mask the reference trains with)")
        self.reserve(max_clips=B, max_tokens=max(Ss, St))
        s = self._f32c(src.detach(), device)
        t = self._f32c(tgt.detach(), device)
        if pe_index is not None:
            pe_index = pe_index.to(device=device, dtype=torch.int32).contiguous()
            if pe_index.numel() != B:
                raise RuntimeError("pe_index must have one entry per clip")
        params = [p for k, p in self.named_parameters() if k != "learned_tgt"]
        self.dropout_seed = int(getattr(self, "dropout_seed", 0))
        return _TrainForward.apply(self, s, t, pe_index, *params)

    def get_tgt_mask(self, size) -> torch.Tensor:
        """models/transformer.py:70-89: (size,size) float CPU tensor, 0 on/below the diagonal, -inf above."""
        mask = torch.tril(torch.ones(size, size) == 1).float()
        mask = mask.masked_fill(mask == 0, float("-inf"))
        mask = mask.masked_fill(mask == 1, float(0.0))
        return mask

    def create_pad_mask(self, matrix: torch.Tensor, pad_token: int) -> torch.Tensor:
        """models/transformer.py:91-93."""
        return matrix == pad_token

    # ------------------------------------------------------------------ rollout (prediction/predict.py)
    def rollout(self, ctx, n_pred, window=5, *, faithful=False, residual=False, teacher=None, pe_index=None,
                scale_in=1.0, scale_out=1.0, out=None):
        """Batched autoregressive rollout on the device: ctx (B,C,E) -> (B,n_pred,E).  See sdvg_rollout.

        pe_index: the positional row of each clip (the reference indexes its table by BATCH position,
        models/positional_encoding.py:35).  None = ``b mod 64``: the clips as one reference batch run in chunks of 64.
        An int = that row for every clip: ``pe_index=0`` reproduces prediction/predict.py, whose DataLoader has
        ``batch_size=1`` (:58), so every clip it rolls out is batch position 0.  Or an int32 tensor (B,)."""
        self._check_eval()
        if ctx.dim() != 3 or ctx.size(2) != self.latent_dim:
            raise RuntimeError(f"expected ctx (B,C,{self.latent_dim}), got {tuple(ctx.shape)}")
        device = ctx.device
        B, Cn = ctx.size(0), ctx.size(1)
        hist = Cn + n_pred
        self.reserve(max_clips=B, max_tokens=6 if faithful else min(window, hist), max_history=hist)
        h = self.engine(device)
        c = self._f32c(ctx, device)
        tptr = None
        if teacher is not None:
            teacher = self._f32c(teacher, device)
            if tuple(teacher.shape) != (B, n_pred, self.latent_dim):
                raise RuntimeError("teacher must be (B, n_pred, E)")
            tptr = teacher.data_ptr()
        pe_ptr = None
        if isinstance(pe_index, int):
            if not 0 <= pe_index < 64:
                raise RuntimeError("pe_index must be a row of the 64-row positional table")
            pe_index = torch.full((B,), pe_index, dtype=torch.int32, device=device)
        if pe_index is not None:
            pe_index = pe_index.to(device=device, dtype=torch.int32).contiguous()
            if pe_index.numel() != B:
                raise RuntimeError("pe_index must have one entry per clip")
            pe_ptr = pe_index.data_ptr()
        if out is None:
            out = torch.empty(B, n_pred, self.latent_dim, device=device, dtype=torch.float32)
        elif out.device != device or out.dtype != torch.float32 or not out.is_contiguous() or \
                tuple(out.shape) != (B, n_pred, self.latent_dim):
            raise RuntimeError("out must be a contiguous fp32 (B, n_pred, E) tensor on the model's device")
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.load().sdvg_rollout(h, c.data_ptr(), B, Cn, n_pred, window,
                                            (1 if faithful else 0) | (2 if residual else 0), tptr, pe_ptr,
                                            float(scale_in), float(scale_out), out.data_ptr(), C.c_void_p(stream)), h)
        return out

    # ------------------------------------------------------------------ instrumentation
    def timing(self, on):
        if self._handle is None:
            raise RuntimeError("no engine yet: run a forward/rollout first")
        _lib.check(_lib.load().sdvg_timing_enable(self._handle, 1 if on else 0), self._handle)

    def timing_read(self):
        n = len(_lib.KERNEL_CLASSES)
        ms, cnt, fl, by = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
        _lib.check(_lib.load().sdvg_timing_read(self._handle, ms, cnt, fl, by), self._handle)
        return {name: dict(ms=ms[i], launches=cnt[i], flops=fl[i], bytes=by[i])
                for i, name in enumerate(_lib.KERNEL_CLASSES)}

    def launch_count(self):
        return 0 if self._handle is None else int(_lib.load().sdvg_launch_count(self._handle))

    def workspace_bytes(self):
        n = C.c_size_t()
        _lib.check(_lib.load().sdvg_workspace_bytes(self._handle, C.byref(n)), self._handle)
        return n.value


class TransformerFuture(Transformer):
    """Drop-in for models/transformer_future.py:9-94: the same network plus the (unused in forward) parameter
    ``learned_tgt`` of shape (1, FRAMES_TO_PREDICT, E) (transformer_future.py:46-47), so its checkpoints load.
    Used one-shot with ``tgt_mask=None`` (prediction/predict_future.py:24,156)."""

    def __init__(self, num_tokens=0, dim_model=256, num_heads=8, num_encoder_layers=6, num_decoder_layers=6,
                 dropout_p=0.1, *, frames_to_predict=None, **kw):
        super().__init__(num_tokens, dim_model, num_heads, num_encoder_layers, num_decoder_layers, dropout_p, **kw)
        if frames_to_predict is None:
            if self.config is None:
                raise ValueError("TransformerFuture(frame_size=...) needs frames_to_predict= (without frame_size it is read "
                                 "from the yaml config like the reference, models/transformer_future.py:46)")
            frames_to_predict = self.config.FRAMES_TO_PREDICT[0]
        self.learned_tgt = nn.Parameter(torch.randn((1, frames_to_predict, self.latent_dim), dtype=torch.float32),
                                        requires_grad=True)

    def _stamp(self):
        return tuple(x for x in super()._stamp() if x[0] != "learned_tgt")

    def state_dict_for_engine(self):
        return {k: v for k, v in self.state_dict().items() if k != "learned_tgt"}


class Identity(nn.Module):
    """Drop-in for models/identity.py:9-41, the copy-last-frame baseline: forward returns ``src[:, -1:]``."""

    def __init__(self):
        super().__init__()

    def forward(self, src, tgt, tgt_mask=None, src_pad_mask=None, tgt_pad_mask=None):
        return src[:, -1:]

    def get_tgt_mask(self, size) -> torch.Tensor:
        mask = torch.tril(torch.ones(size, size) == 1).float()
        mask = mask.masked_fill(mask == 0, float("-inf"))
        return mask.masked_fill(mask == 1, float(0.0))

    def create_pad_mask(self, matrix: torch.Tensor, pad_token: int) -> torch.Tensor:
        return matrix == pad_token
