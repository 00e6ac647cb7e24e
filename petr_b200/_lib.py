"""ctypes binding of libpetr_b200.so (the C ABI declared in include/petr_b200.h).

The library is the product: there is no Python/PyTorch fallback for any op.  If the shared
object is missing or an entry point fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpetr_b200.so")

_P, _I, _F = c_void_p, c_int, c_float
ABI_VERSION = 3

# name -> (restype, argtypes); mirrors include/petr_b200.h one to one
SIGNATURES = {
    "petr_abi_version": (c_int, []),
    "petr_last_error_string": (c_char_p, []),
    "petr_launch_count": (c_int64, []),
    "petr_launch_count_reset": (None, []),
    "petr_set_tuning": (c_int, [c_char_p, _I]),
    "petr_get_tuning": (c_int, [c_char_p]),
    "petr_debug_set_trace": (None, [_P]),
    "petr_key_padding_mask": (c_int, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "petr_frustum_features": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P, _P,
                                      _P, _I, _P, _P, _P]),
    "petr_sine3d_features": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _P, _I, _P]),
    "petr_pos2posemb3d": (c_int, [_P, _P, _I, _I, _P, _P]),
    "petr_gemm_f32": (c_int, [_P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P]),
    "petr_attention_f32": (c_int, [_P, _I, _P, _I, _P, _I, _P, _P, _I, _P, _I, _P, _P, _F, c_uint64, _P, _I,
                                   _I, _I, _I, _I, _I, _F, _P]),
    "petr_layernorm_multi_f32": (c_int, [_P, _P, _P, _F, _P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, c_int64, _P]),
    "petr_split_rows_bf16": (c_int, [_P, _I, _P, _I, _P, _I, _I, _I, _P]),
    "petr_x3_weight_bf16": (c_int, [_P, _I, _I, _I, _P, _P]),
    "petr_nchw_to_tokens_bf16": (c_int, [_P, _I, _P, _I, _I, _I, _P]),
    "petr_layernorm_f32": (c_int, [_P, _P, _P, _P, _I, _I, _F, _I, _I, _P]),
    "petr_nms_free_decode": (c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _I, _F, _I, _P, _P, _P, _P, _P, _P, _P]),
    "petr_decode_boxes": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "petr_gemm_bf16_tc": (c_int, [_P, _I, _I, _P, _I, _P, _P, _P, _I, _F, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "petr_gemm_tf32x3_tc": (c_int, [_P, _I, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "petr_split_rows_tf32": (c_int, [_P, _I, _P, _I, _P, _I, _I, _P]),
    "petr_position_mlp_fused": (c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P,
                                        _I, _I, _P, _P, _I, _P, _P, _P]),
    "petr_stage_small": (c_int, [_P, _P, _I, _P]),
    "petr_add_rows_bf16": (c_int, [_P, _P, _I, _P, _I, _I, _P]),
    "petr_pack_q_bf16": (c_int, [_P, _I, _P, _I, _I, _I, _F, _P]),
    "petr_cross_attention_bf16": (c_int, [_P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "petr_cross_attention_workspace_bytes": (c_int64, [_I, _I, _I, _I, _I]),
    "petr_cross_attention_trace": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "petr_attention_x3_tc": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _F, _I, _I, _P]),
    "petr_attention_x3_workspace_bytes": (c_int64, [_I, _I, _I, _I, _I]),
    "petr_attention_merge": (c_int, [_P, _I, _P, _I, _P, _P, _I, _I, _I, _P]),
    "petr_attention_partial_floats": (c_int64, [_I, _I, _I]),
    "petr_sine2d_features": (c_int, [_P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _P, _I, _P]),
    "petr_learned_pos_features": (c_int, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    "petr_tokens_to_nchw_f32": (c_int, [_P, _P, _I, _I, _I, _P]),
    "petr_upsample_add_tokens": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "petr_im2col3x3_tokens": (c_int, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "petr_norm_scale_f32": (c_int, [_P, _P, _I, _I, _I, _I, _F, _F, _F, _P]),
    "petr_p2p_push_partial": (c_int, [_P, _P, _P, _I, _I, c_int64, _I, _P, _I, _I, _P, _P]),
    "petr_p2p_epoch_inc": (c_int, [_P, _P]),
    "petr_attention_merge_p2p": (c_int, [_P, _I, _I, _P, _P, _I, _I, _P, _P, _I, _P, _I, _I, _I, _P]),
    "petr_gemm_f32_tn": (c_int, [_P, _I, _P, _I, _I, _I, _P, _I, _P, _I, _I, _I, _P]),
    "petr_split_transpose_bf16": (c_int, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    "petr_sum_parts_f32": (c_int, [_P, _I, c_int64, _P, _P]),
    "petr_colsum_f32": (c_int, [_P, _I, _I, _I, _P, _P]),
    "petr_rowsum_periodic_f32": (c_int, [_P, _I, _I, _I, _P, _P]),
    "petr_relu_bwd_f32": (c_int, [_P, _P, _P, c_int64, _P]),
    "petr_sigmoid_gate_bwd_f32": (c_int, [_P, _P, _P, _P, _P, c_int64, _P]),
    "petr_dropout_f32": (c_int, [_P, _P, _P, c_int64, _F, c_uint64, _P]),
    "petr_norm_scale_bwd_f32": (c_int, [_P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _P]),
    "petr_learned_pos_bwd": (c_int, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "petr_layernorm_bwd_f32": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _I, _I, _P]),
    "petr_attention_f32_bwd": (c_int, [_P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _P, _P, _I,
                                       _P, _I, _P, _I, _P, _I, _F, c_uint64, _I, _I, _I, _I, _I, _F, _P]),
    "petr_pos2posemb3d_bwd": (c_int, [_P, _P, _P, _I, _I, _P, _P]),
    "petr_decode_boxes_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "petr_normalize_bbox": (c_int, [_P, _I, _I, _P, _P]),
    "petr_hungarian_cost": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _F, _F, _P, _P]),
    "petr_focal_loss": (c_int, [_P, _P, _P, c_int64, _I, _F, _F, _F, _P, _P, _P]),
    "petr_l1_loss": (c_int, [_P, _P, _P, c_int64, _I, _F, _P, _P, _P]),
}

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m petr_b200.build` "
                "(petr_b200 has no CPU / PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.petr_abi_version() != ABI_VERSION:
            raise RuntimeError("libpetr_b200.so ABI version mismatch")
        _lib = L
    return _lib


def ptr(t):
    """Device (or host) pointer of a tensor / None."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def host_floats(vals):
    return (c_float * len(vals))(*[float(v) for v in vals])


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def check(rc: int, name: str):
    if rc != 0:
        msg = lib().petr_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{name} failed with code {rc}: {msg}")


# Optional live timing (bench.py roofline / --breakdown): when TIMERS is a dict, every C-ABI call
# is bracketed by CUDA events recorded on torch's current stream; key = LABEL (set by the ops
# wrapper for named GEMMs) or the entry-point name.
TIMERS = None
LABEL = None


def call(name: str, *args):
    if TIMERS is None:
        check(getattr(lib(), name)(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib(), name)(*args), name)
    e1.record()
    TIMERS.setdefault(LABEL or name, []).append((e0, e1))


def set_tuning(key: str, value: int):
    check(lib().petr_set_tuning(key.encode(), int(value)), "petr_set_tuning")


def get_tuning(key: str) -> int:
    return int(lib().petr_get_tuning(key.encode()))


def launch_count() -> int:
    return int(lib().petr_launch_count())


def reset_launch_count():
    lib().petr_launch_count_reset()
