"""ctypes binding of ``libmicformer_b200.so`` (the C ABI declared in ``include/micformer_b200.h``).

There is deliberately NO fallback: if the library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libmicformer_b200.so")

P, I, L, F, D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> argument ctypes (return type is int unless listed in _RESTYPES)
SIGNATURES = {
    "mic_version": [],
    "mic_last_error_string": [],
    "mic_launch_count": [],
    "mic_reset_launch_count": [],
    "mic_set_gemm_mode": [I],
    "mic_get_gemm_mode": [],
    "mic_layernorm_fwd": [P, I, P, I, P, P, P, P, P, I, I, I, I, I, I, I, F, P],
    "mic_layernorm_bwd": [P, P, I, P, I, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, P],
    "mic_linear_fwd": [P, I, P, I, I, P, P, I, I, I, I, I, P, I, P, I, P, I, I, P],
    "mic_linear_bwd_data": [P, I, P, I, I, P, I, I, I, I, P, I, P, I, I, P],
    "mic_linear_bwd_weight": [P, I, P, I, P, I, I, P, I, I, I, P, I, P],
    "mic_window_attn_fwd": [P, I, P, P, I, P, I, P, I, I, I, I, I, I, I, I, I, F, P],
    "mic_window_attn_bwd": [P, I, P, P, I, P, P, I, P, P, I, P, P, I, I, I, I, I, I, I, I, I, I, F, P],
    "mic_conv3_fwd": [P, I, P, I, P, P, P, I, I, I, I, I, I, I, I, I, P],
    "mic_conv3_bwd_data": [P, P, P, I, I, P, I, I, I, I, I, I, I, I, I, I, I, P],
    "mic_conv3_bwd_weight": [P, P, I, P, I, P, P, I, I, I, I, I, I, I, I, I, P],
    "mic_offset_head_fwd": [P, P, P, P, P, I, I, I, I, I, F, P],
    "mic_offset_head_bwd": [P, P, P, P, P, P, P, P, P, I, I, I, I, I, F, P],
    "mic_deform_sample_fwd": [P, P, P, I, I, I, I, I, I, I, I, P],
    "mic_deform_sample_bwd": [P, P, P, P, P, I, I, I, I, I, I, I, I, P],
    "mic_block_permute": [P, P, I, I, I, I, I, I, L, I, P],
    "mic_dice_bce_partial": [P, P, P, I, I, L, P],
    "mic_dice_bce_finalize": [P, P, P, I, D, P],
    "mic_dice_bce_bwd": [P, P, P, P, P, I, I, L, D, P],
    "mic_crop_residual": [P, P, P, P, I, I, I, I, I, I, I, I, P],
    "mic_crop_residual_bwd": [P, P, P, I, I, I, I, I, I, I, I, P],
}
_RESTYPES = {"mic_last_error_string": C.c_char_p, "mic_launch_count": C.c_int64, "mic_reset_launch_count": None}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built -- there is no CPU/torch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"micformer_b200: native library not found at {LIB_PATH}. Build it with "
            f"`python -m micformer_b200.build` (nvcc, sm_100a). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing: loud by design
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def last_error() -> str:
    return load().mic_last_error_string().decode(errors="replace")


def ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    """Invoke an ``int``-returning entry point on the current CUDA stream; raise on a non-zero code."""
    lib = load()
    rc = getattr(lib, name)(*args, stream_ptr())
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error()}")


def check_cuda_f32(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("micformer_b200 runs on CUDA tensors only (sm_100a kernels; no CPU fallback)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"micformer_b200 kernels take float32 tensors, got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError("micformer_b200 kernels take contiguous tensors")


def launch_count() -> int:
    return int(load().mic_launch_count())


def reset_launch_count() -> None:
    load().mic_reset_launch_count()


def set_gemm_mode(mode: int) -> None:
    rc = load().mic_set_gemm_mode(int(mode))
    if rc != 0:
        raise RuntimeError(last_error())


def get_gemm_mode() -> int:
    return int(load().mic_get_gemm_mode())
