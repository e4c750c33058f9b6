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
    "mic_linear_unpatch_view": [I, I, I, I],
    "mic_window_attn_fwd": [P, I, P, P, I, P, I, P, I, I, I, I, I, I, I, I, I, F, P],
    "mic_window_attn_bwd": [P, I, P, P, I, P, P, I, P, P, I, P, P, I, I, I, I, I, I, I, I, I, I, F, P],
    "mic_conv3_fwd": [P, I, P, I, P, P, P, I, I, I, I, I, I, I, I, I, P],
    "mic_conv3_tc_fwd": [P, I, P, I, P, P, P, I, I, I, I, I, I, P],
    "mic_conv3_bwd_data": [P, P, P, I, I, P, I, I, I, I, I, I, I, I, I, I, I, P],
    "mic_conv3_tc_bwd_data": [P, P, P, I, I, P, I, I, I, I, I, I, I, I, P],
    "mic_conv3_bwd_weight": [P, P, I, P, I, P, P, I, I, I, I, I, I, I, I, I, P],
    "mic_conv3_mma_bwd_weight": [P, P, I, P, I, P, P, I, I, I, I, I, I, I, P],
    "mic_offset_head_fwd": [P, P, P, P, P, I, I, I, I, I, F, P],
    "mic_offset_head_bwd": [P, P, P, P, P, P, P, P, P, I, I, I, I, I, F, P],
    "mic_deform_sample_fwd": [P, P, P, I, I, I, I, I, I, I, I, P],
    "mic_deform_sample_bwd": [P, P, P, P, P, I, I, I, I, I, I, I, I, P],
    "mic_block_permute": [P, P, I, I, I, I, I, I, L, I, P],
    "mic_dice_bce_partial": [P, P, P, I, I, L, P],
    "mic_dice_bce_finalize": [P, P, P, I, D, P],
    "mic_dice_bce_bwd": [P, P, P, P, P, I, I, L, D, P],
    "mic_dice_bce_partial_u8": [P, P, P, I, I, L, P],
    "mic_dice_bce_bwd_u8": [P, P, P, P, P, I, I, L, P],
    "mic_dice_bce_finalize_weighted": [P, P, P, I, D, D, D, P],
    "mic_adam_chunk_elems": [],
    "mic_adam_step": [P, P, P, P, P, P, P, I, I, P, P, F, F, F, F, P],
    "mic_weight_images": [P, I, L, P],
    "mic_conv_weight_layouts": [P, I, L, I, P],
    "mic_mlp_block_fwd": [P, P, P, P, P, P, P, P, P, P, P, I, I, I, F, P],
    "mic_mlp_block_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, P, I, P, P, P, P, P, P, I, I, F, P],
    "mic_mlp_block_smem": [I],
    "mic_debug_t5_trace": [P],
    "mic_mlp_split_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, I, P, P, P, P, I, I, I, F, P],
    "mic_mlp_split_fwd": [P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, F, P],
    "mic_attn_block_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, F, F, P],
    "mic_attn_block_fwd": [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, F, F, P],
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
    mode = os.environ.get("MICFORMER_GEMM_MODE")
    if mode is not None:
        if lib.mic_set_gemm_mode(int(mode)) != 0:
            raise RuntimeError(f"MICFORMER_GEMM_MODE={mode}: " + lib.mic_last_error_string().decode())
    return lib


def unpatch_view(ch: int, dc: int, hc: int, wc: int) -> None:
    """The next mic_linear_* call of this thread addresses its Y / dY matrix through the fine channels-last grid
    (include/micformer_b200.h: mic_linear_unpatch_view)."""
    if load().mic_linear_unpatch_view(ch, dc, hc, wc) != 0:
        raise RuntimeError("mic_linear_unpatch_view: " + last_error())


def last_error() -> str:
    return load().mic_last_error_string().decode(errors="replace")


def ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


_prof = None   # name -> [events..., flops, bytes, count] when profiling is on (bench.py's per-kernel pass)


def _prod(*v):
    r = 1
    for x in v:
        r *= x
    return r


# algorithmic (flops, bytes) per call from the C-ABI arguments (stream excluded); DESIGN.md states the formulas
def _cost_conv3(B, D, H, W, Dp, Hp, Wp, Cin, Co):
    P = B * Dp * Hp * Wp
    return 2 * 27 * Cin * Co * P, 4 * (B * D * H * W * Cin + P * Co + 27 * Cin * Co)


COST = {
    "mic_linear_fwd": lambda a: (2 * a[8] * a[9] * a[10], 4 * (a[8] * a[10] + a[9] * a[10] + a[8] * a[9] *
                                                                (1 + (a[12] is not None) + (a[14] is not None)))),
    "mic_linear_bwd_data": lambda a: (2 * a[7] * a[8] * a[9], 4 * (a[7] * a[8] + a[8] * a[9] + a[7] * a[9] *
                                                                   (1 + (a[10] is not None)))),
    "mic_linear_bwd_weight": lambda a: (2 * a[8] * a[9] * a[10], 4 * (a[8] * a[9] + a[8] * a[10] + a[9] * a[10])),
    "mic_conv3_fwd": lambda a: _cost_conv3(a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[1] + a[3], a[14]),
    "mic_conv3_tc_fwd": lambda a: _cost_conv3(a[7], a[8], a[9], a[10], a[8], a[9], a[10], a[1] + a[3], a[11]),
    "mic_conv3_bwd_data": lambda a: _cost_conv3(a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[3] + a[6], a[15]),
    "mic_conv3_tc_bwd_data": lambda a: _cost_conv3(a[8], a[9], a[10], a[11], a[9], a[10], a[11], a[3] + a[6], a[12]),
    "mic_conv3_bwd_weight": lambda a: _cost_conv3(a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[2] + a[4], a[14]),
    "mic_conv3_mma_bwd_weight": lambda a: _cost_conv3(a[7], a[8], a[9], a[10], a[8], a[9], a[10], a[2] + a[4], a[11]),
    "mic_window_attn_fwd": lambda a: (4 * _prod(*a[8:12]) * a[12] * a[13] * _prod(*a[14:17]),
                                      4 * (4 * _prod(*a[8:12]) * a[12] * a[13] + _prod(*a[8:12]) * a[12])),
    "mic_window_attn_bwd": lambda a: (10 * _prod(*a[14:18]) * a[18] * a[19] * _prod(*a[20:23]),
                                      4 * (8 * _prod(*a[14:18]) * a[18] * a[19] + _prod(*a[14:18]) * a[18])),
    "mic_layernorm_fwd": lambda a: (0, 4 * (a[1] + a[3]) * (_prod(*a[9:13]) + a[9] * _prod(*a[13:16]))),
    "mic_layernorm_bwd": lambda a: (0, 4 * (a[2] + a[4]) * (2 * _prod(*a[14:18]) + a[14] * _prod(*a[18:21]) +
                                                          (a[8] is not None) * _prod(*a[14:18]))),
    "mic_deform_sample_fwd": lambda a: (0, 4 * (a[3] * _prod(*a[4:7]) * a[10] + a[3] * _prod(*a[7:10]) * (a[10] + 3))),
    "mic_deform_sample_bwd": lambda a: (0, 4 * (2 * a[5] * _prod(*a[6:9]) * a[12] + a[5] * _prod(*a[9:12]) * (a[12] + 6))),
    "mic_block_permute": lambda a: (0, 8 * a[2] * a[3] * a[4] * a[5] * a[6] ** 3 * a[7]),
    "mic_dice_bce_partial": lambda a: (0, 8 * a[3] * a[4] * a[5]),
    "mic_dice_bce_bwd": lambda a: (0, 12 * a[5] * a[6] * a[7]),
    "mic_dice_bce_partial_u8": lambda a: (0, 5 * a[3] * a[4] * a[5]),
    "mic_dice_bce_bwd_u8": lambda a: (0, 9 * a[5] * a[6] * a[7]),
    "mic_offset_head_fwd": lambda a: (0, 4 * _prod(*a[5:9]) * (a[9] + 3)),
    "mic_offset_head_bwd": lambda a: (0, 4 * _prod(*a[9:13]) * (2 * a[13] + 3)),
    "mic_mlp_block_fwd": lambda a: (16 * a[12] * a[13] * a[13], 8 * a[12] * a[13] + 4 * 8 * a[13] * a[13]),
    "mic_mlp_block_bwd": lambda a: (40 * a[20] * a[21] * a[21], 12 * a[20] * a[21] + 4 * 16 * a[21] * a[21]),
    "mic_attn_block_fwd": lambda a: (_prod(*a[15:19]) * (8 * a[19] * a[19] + 32 * a[19]),
                                     4 * _prod(*a[15:19]) * a[19] * (3 if a[1] is not None else 2) + 32 * a[19] * a[19]),
    "mic_attn_block_bwd": lambda a: (_prod(*a[19:23]) * (24 * a[23] * a[23] + 80 * a[23]),
                                     4 * _prod(*a[19:23]) * a[23] * (5 if a[1] is not None else 3) + 64 * a[23] * a[23]),
    "mic_mlp_split_fwd": lambda a: (16 * a[12] * a[13] * a[13], 8 * a[12] * a[13] + 4 * 8 * a[13] * a[13]),
    "mic_mlp_split_bwd": lambda a: (40 * a[20] * a[21] * a[21], 12 * a[20] * a[21] + 4 * 16 * a[21] * a[21]),
    "mic_crop_residual": lambda a: (0, 12 * _prod(*a[4:8]) * a[11]),
    "mic_crop_residual_bwd": lambda a: (0, 4 * a[3] * (_prod(*a[4:7]) + _prod(*a[7:10])) * a[10]),
}


TAG = {
    "mic_mlp_split_bwd": lambda a: f"T{a[20]}xC{a[21]}",
    "mic_mlp_split_fwd": lambda a: f"T{a[12]}xC{a[13]}",
    "mic_attn_block_bwd": lambda a: f"T{a[19] * a[20] * a[21] * a[22]}xC{a[23]}",
    "mic_attn_block_fwd": lambda a: f"T{a[15] * a[16] * a[17] * a[18]}xC{a[19]}",
    "mic_mlp_block_fwd": lambda a: f"T{a[12]}xC{a[13]}",
    "mic_mlp_block_bwd": lambda a: f"T{a[20]}xC{a[21]}",
    "mic_linear_fwd": lambda a: f"M{a[8]}xN{a[9]}xK{a[10]}",
    "mic_linear_bwd_data": lambda a: f"M{a[7]}xN{a[8]}xK{a[9]}",
    "mic_linear_bwd_weight": lambda a: f"M{a[8]}xN{a[9]}xK{a[10]}",
    "mic_conv3_fwd": lambda a: f"Cin{a[1] + a[3]}xCo{a[14]}@{a[7] * a[11] * a[12] * a[13]}",
    "mic_conv3_tc_fwd": lambda a: f"Cin{a[1] + a[3]}xCo{a[11]}@{a[7] * a[8] * a[9] * a[10]}",
    "mic_conv3_bwd_data": lambda a: f"Cin{a[3] + a[6]}xCo{a[15]}@{a[8] * a[12] * a[13] * a[14]}",
    "mic_conv3_tc_bwd_data": lambda a: f"Cin{a[3] + a[6]}xCo{a[12]}@{a[8] * a[9] * a[10] * a[11]}",
    "mic_conv3_bwd_weight": lambda a: f"Cin{a[2] + a[4]}xCo{a[14]}@{a[7] * a[11] * a[12] * a[13]}",
    "mic_conv3_mma_bwd_weight": lambda a: f"Cin{a[2] + a[4]}xCo{a[11]}@{a[7] * a[8] * a[9] * a[10]}",
}


def profile_begin() -> None:
    global _prof
    _prof = {}


def profile_end() -> dict:
    """-> {name: {"ms": total device ms, "calls": n, "flops": algorithmic flops, "bytes": algorithmic bytes}}"""
    global _prof
    torch.cuda.synchronize()
    prof, _prof = (_prof or {}), None
    # calibrate: event-pair overhead around the smallest kernel of the library (1 thread of work)
    lib = load()
    d = torch.zeros(8, dtype=torch.float64, device="cuda"); f = torch.zeros(8, device="cuda")
    over = []
    for _ in range(20):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.mic_dice_bce_finalize(d.data_ptr(), f.data_ptr(), f.data_ptr() + 4, 1, 1.0, stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        over.append(e0.elapsed_time(e1))
    over.sort()
    base = max(0.0, over[len(over) // 2] - 0.002)       # that kernel itself runs ~2 us
    out = {}
    for name, rec in prof.items():
        ms = sum(max(e0.elapsed_time(e1) - base, 0.0005) for e0, e1 in rec["ev"])
        out[name] = {"ms": ms, "calls": len(rec["ev"]), "flops": rec["flops"], "bytes": rec["bytes"]}
    return out


def _invoke(name: str, args, allow_unsupported: bool) -> bool:
    lib = load()
    if _prof is not None:
        # the device is drained first so the event pair brackets only this launch (not host queueing gaps);
        # profile_end() subtracts the calibrated empty-launch overhead
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream_ptr())
        e1.record()
        if rc == 0:
            tg = TAG.get(name)
            key = f"{name}[{tg(args)}]" if tg is not None else name
            rec = _prof.setdefault(key, {"ev": [], "flops": 0, "bytes": 0})
            rec["ev"].append((e0, e1))
            fn = COST.get(name)
            if fn is not None:
                f, b = fn(args)
                rec["flops"] += f; rec["bytes"] += b
    else:
        rc = getattr(lib, name)(*args, stream_ptr())
    if rc == -2 and allow_unsupported:
        return False
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error()}")
    return True


def call(name: str, *args):
    """Invoke an ``int``-returning entry point on the current CUDA stream; raise on a non-zero code."""
    _invoke(name, args, False)


_declined = set()


def try_call(name: str, *args) -> bool:
    """Like ``call`` but returns False when the entry point declines the shape (MIC_ERR_UNSUPPORTED); the caller then
    runs the exact fp32 CUDA-core kernel.  Reported once per entry point (MICFORMER_WARN_FALLBACK=0 silences it)."""
    ok = _invoke(name, args, True)
    if not ok and name not in _declined:
        _declined.add(name)
        if os.environ.get("MICFORMER_WARN_FALLBACK", "1") != "0":
            import warnings
            warnings.warn(f"micformer_b200: {name} declined a shape ({last_error()}); using the fp32 CUDA-core kernel "
                          f"(reported once per entry point)", RuntimeWarning, stacklevel=3)
    return ok


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
