"""ctypes binding of libsdvg.so (include/sdvg.h).  No torch types cross this boundary: only raw device
pointers, sizes and the CUDA stream handle.  There is no CPU fallback: if the shared library cannot be
loaded, or no sm_100 device is present, the errors propagate as RuntimeError."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsdvg.so")

PRECISIONS = {"fp32_simt": 0, "fp32": 1, "fp16": 2, "bf16": 3, "mixed": 4}
KERNEL_CLASSES = ("gemm_tc", "gemm_simt", "attention", "layernorm", "pack", "persistent")

# every symbol include/sdvg.h declares (tests check the library exports all of them)
SYMBOLS = (
    "sdvg_version", "sdvg_last_error", "sdvg_create", "sdvg_destroy", "sdvg_workspace_bytes", "sdvg_set_weight",
    "sdvg_num_weights", "sdvg_weight_key", "sdvg_finalize_weights", "sdvg_forward", "sdvg_rollout",
    "sdvg_timing_enable", "sdvg_timing_read", "sdvg_launch_count", "sdvg_gemm", "sdvg_criterion",
    "sdvg_train_backward", "sdvg_train_gradients", "sdvg_train_set_ready_callback", "sdvg_train_set_dropout", "sdvg_param_range", "sdvg_train_prediction", "sdvg_train_adam_step", "sdvg_train_adam_step_range",
    "sdvg_get_weight", "sdvg_train_forward", "sdvg_train_backward_from",
)


class SdvgConfig(C.Structure):
    _fields_ = [("dim_model", C.c_int32), ("num_heads", C.c_int32), ("num_encoder_layers", C.c_int32),
                ("num_decoder_layers", C.c_int32), ("latent_dim", C.c_int32), ("dim_feedforward", C.c_int32),
                ("layer_norm_eps", C.c_float), ("max_clips", C.c_int32), ("max_tokens", C.c_int32),
                ("max_history", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32)]


class SdvgLossConfig(C.Structure):
    _fields_ = [("frames_to_predict", C.c_int32), ("use_mse", C.c_int32), ("use_l1", C.c_int32), ("use_gdl", C.c_int32),
                ("lambda_gdl", C.c_float), ("alpha", C.c_float), ("use_contrastive", C.c_int32),
                ("temperature", C.c_float), ("lambda_contrastive", C.c_float)]


GRAD_READY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_longlong, C.c_longlong)

_lib = None


def load(build_if_missing=True):
    """Load (building first if the in-tree .so is missing or stale and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        try:
            from . import build as _build
            _build.build()
        except FileNotFoundError:
            # no nvcc on this machine: use the shipped library if there is one (compile errors are NOT swallowed)
            if not os.path.exists(LIB_PATH):
                raise
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python sd-video-gen_b200/build.py` "
                           "(libsdvg has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
    lib.sdvg_version.restype = C.c_int
    lib.sdvg_last_error.restype = C.c_char_p
    lib.sdvg_last_error.argtypes = [vp]
    lib.sdvg_create.argtypes = [C.POINTER(SdvgConfig), C.POINTER(vp)]
    lib.sdvg_destroy.argtypes = [vp]
    lib.sdvg_destroy.restype = None
    lib.sdvg_workspace_bytes.argtypes = [vp, C.POINTER(C.c_size_t)]
    lib.sdvg_set_weight.argtypes = [vp, C.c_char_p, vp, C.POINTER(C.c_int64), i32]
    lib.sdvg_num_weights.argtypes = [vp]
    lib.sdvg_weight_key.argtypes = [vp, i32]
    lib.sdvg_weight_key.restype = C.c_char_p
    lib.sdvg_finalize_weights.argtypes = [vp, vp]
    lib.sdvg_forward.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.sdvg_rollout.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, f32, f32, vp, vp]
    lib.sdvg_timing_enable.argtypes = [vp, i32]
    lib.sdvg_timing_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]
    lib.sdvg_launch_count.argtypes = [vp]
    lib.sdvg_launch_count.restype = C.c_int64
    lib.sdvg_gemm.argtypes = [i32, i32, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, C.POINTER(f32), vp]
    lib.sdvg_criterion.argtypes = [i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, f32, i32, f32, f32, vp, vp]
    lib.sdvg_train_backward.argtypes = [vp, vp, vp, vp, i32, i32, i32, C.POINTER(SdvgLossConfig), vp, vp, i32, vp]
    lib.sdvg_train_gradients.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.sdvg_train_set_ready_callback.argtypes = [vp, GRAD_READY_FN, vp, i32]
    lib.sdvg_train_set_dropout.argtypes = [vp, f32, C.c_uint64]
    lib.sdvg_param_range.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.sdvg_train_prediction.argtypes = [vp, C.POINTER(vp)]
    lib.sdvg_train_adam_step.argtypes = [vp, f32, f32, f32, f32, f32, vp]
    lib.sdvg_train_adam_step_range.argtypes = [vp, f32, f32, f32, f32, f32, C.c_int64, C.c_int64, i32, vp]
    lib.sdvg_get_weight.argtypes = [vp, C.c_char_p, vp, vp]
    lib.sdvg_train_forward.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp]
    lib.sdvg_train_backward_from.argtypes = [vp, vp, vp]
    _lib = lib
    return lib


def check(rc, handle=None):
    if rc != 0:
        msg = load().sdvg_last_error(handle)
        raise RuntimeError(f"libsdvg error {rc}: {msg.decode() if msg else '?'}")
