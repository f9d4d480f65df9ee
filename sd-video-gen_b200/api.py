"""Public surface of the package (imported as ``sdvg_b200``)."""
from . import _lib
from .config import CONFIGS, latent_dim, parse_config_args
from .dist import pe_index_for, rollout_sharded, shard_bounds
from .latent_cache import load_latent_clips, load_latent_frames, save_latent_frames
from .positional_encoding import PositionalEncoding
from .predict import (LATENT_SCALE, SOS_VALUE, HostRollout, predict, predict_diff, predict_future, rollout,
                      rollout_from_host)
from .trainer import (AdamTrainer, BiPatchNCE, allreduce_buckets, criterion, gradient_difference_loss, loss_terms,
                      validation_step)
from .transformer import Identity, Transformer, TransformerFuture

__all__ = ["Transformer", "TransformerFuture", "Identity", "PositionalEncoding", "predict", "predict_diff",
           "predict_future", "rollout", "rollout_from_host", "HostRollout",
           "rollout_sharded", "shard_bounds", "pe_index_for", "CONFIGS", "latent_dim", "parse_config_args",
           "LATENT_SCALE", "SOS_VALUE", "gemm", "build_library", "criterion", "gradient_difference_loss", "BiPatchNCE",
           "loss_terms", "validation_step", "AdamTrainer", "allreduce_buckets", "load_latent_clips", "load_latent_frames",
           "save_latent_frames"]


def build_library(force=False, verbose=False):
    from . import build
    return build.build(force=force, verbose=verbose)


def gemm(A, W, bias=None, relu=False, precision="fp32", block_n=0, iters=1):
    """C = A @ W^T (+bias)(ReLU) through libsdvg's GEMM kernels (unit tests / micro-benchmarks).
    Returns (C, ms_per_launch)."""
    import ctypes as C_
    import torch
    if not A.is_cuda:
        raise RuntimeError("sdvg_b200.gemm needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    A = A.float().contiguous()
    W = W.float().contiguous()
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, device=A.device, dtype=torch.float32)
    b = None if bias is None else bias.float().contiguous()
    ms = C_.c_float(0)
    prec = _lib.PRECISIONS["fp16" if precision == "mixed" else precision]
    stream = torch.cuda.current_stream(A.device).cuda_stream
    rc = lib.sdvg_gemm(A.device.index or 0, prec, A.data_ptr(), W.data_ptr(), None if b is None else b.data_ptr(),
                       1 if relu else 0, out.data_ptr(), M, N, K, block_n, iters, C_.byref(ms), C_.c_void_p(stream))
    _lib.check(rc)
    return out, ms.value
