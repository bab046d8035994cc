"""ctypes loader for libsaspa_b200.so (the C ABI declared in include/saspa_b200.h).

There is no fallback: if the library is missing or an entry point fails, a
``SaspaError`` is raised.  ``load()`` never builds; ``saspa_aug_b200.build.build()``
does (called by ``__graft_entry__.build()``)."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libsaspa_b200.so")


class SaspaError(RuntimeError):
    pass


class Epilogue(Structure):
    """Mirror of ``saspa_epilogue`` (include/saspa_b200.h)."""

    _fields_ = [
        ("bias", c_void_p),
        ("row_bias", c_void_p),
        ("rows_per_group", c_int),
        ("ld_row_bias", c_int),
        ("act", c_int),
        ("alpha", c_float),
        ("residual", c_void_p),
        ("ld_res", c_int),
        ("beta", c_float),
        ("out_fp32", c_int),
        ("act_after_residual", c_int),
        ("row_stats_out", c_void_p),
        ("row_stats_slots", c_int),
        ("ln_stats", c_void_p),
        ("ln_slots", c_int),
        ("ln_colsum", c_void_p),
        ("ln_eps", c_float),
    ]


class LinComb(Structure):
    """Mirror of ``saspa_lincomb``."""

    _fields_ = [
        ("inp", c_void_p * 8),
        ("out", c_void_p * 4),
        ("coef", c_float * 32),
        ("n_in", c_int),
        ("n_out", c_int),
    ]


ACT_NONE, ACT_SILU, ACT_GELU, ACT_RELU, ACT_QUICKGELU, ACT_GEGLU = range(6)

# name -> (restype, argtypes); every symbol include/saspa_b200.h declares
_P = c_void_p
SIGNATURES = {
    "saspa_version": (c_int, []),
    "saspa_last_error_string": (c_char_p, []),
    "saspa_canny_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "saspa_canny_u8": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_size_t, _P]),
    "saspa_pil_coeffs_host": (c_int, [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int32), POINTER(c_int32), c_int]),
    "saspa_pil_ksize": (c_int, [c_int, c_int, c_int]),
    "saspa_resize_pil_u8": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, _P, c_int, _P, _P, c_int, _P]),
    "saspa_crop_normalize_bf16": (
        c_int,
        [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_float, _P, c_int, _P],
    ),
    "saspa_gemm_bf16": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, POINTER(Epilogue), _P]),
    "saspa_gemm_row_stats_slots": (c_int, [c_int]),
    "saspa_conv2d_igemm_bf16": (
        c_int,
        [_P, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, c_int, c_int, POINTER(Epilogue), _P],
    ),
    "saspa_conv2d_igemm_strided_bf16": (
        c_int,
        [_P, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, POINTER(Epilogue), _P],
    ),
    "saspa_im2col_bf16": (
        c_int,
        [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P],
    ),
    "saspa_groupnorm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "saspa_groupnorm_nhwc_bf16": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, c_int, _P, c_int, _P, c_size_t, _P]),
    "saspa_layernorm_bf16": (c_int, [_P, c_int, c_int, c_int, c_float, _P, _P, _P, c_int, _P]),
    "saspa_act_bf16": (c_int, [_P, _P, c_size_t, c_int, _P]),
    "saspa_add_bf16": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, _P]),
    "saspa_upsample_nearest2x_bf16": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "saspa_nchw_f32_to_nhwc_bf16": (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, c_float, _P]),
    "saspa_nhwc_to_nchw_f32": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "saspa_pool2d_nhwc_bf16": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P]),
    "saspa_attention_bf16": (
        c_int,
        [_P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _P],
    ),
    "saspa_softmax_rows_bf16": (c_int, [_P, c_int, _P, c_int, ctypes.c_longlong, c_int, c_float, _P]),
    "saspa_transpose_bf16": (c_int, [_P, c_int, ctypes.c_longlong, _P, c_int, ctypes.c_longlong, c_int, c_int, c_int, _P]),
    "saspa_timestep_sinusoid_bf16": (c_int, [_P, c_int, c_int, c_int, c_float, _P, _P]),
    "saspa_cfg_sched_step": (c_int, [_P, _P, c_float, POINTER(LinComb), c_size_t, _P]),
    "saspa_cfg_sched_step_clip": (c_int, [_P, _P, c_float, POINTER(LinComb), POINTER(c_float), POINTER(c_float), c_float, c_size_t, _P]),
    "saspa_vae_sample_add_noise": (c_int, [_P, _P, _P, c_float, c_float, c_float, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "saspa_vae_quantize_u8": (c_int, [_P, c_int, c_int, c_size_t, _P, _P]),
    "saspa_bap_head": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "saspa_fc_f32": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "saspa_topk_contains": (c_int, [_P, c_int, c_int, _P, c_int, _P, _P, _P]),
    "saspa_softmax_at_f32": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P]),
    "saspa_clip_score_argmax": (c_int, [_P, _P, c_int, c_int, c_int, c_float, _P, _P, _P]),
    "saspa_rgb_to_luma3_u8": (c_int, [_P, ctypes.c_longlong, _P, _P]),
    "saspa_lpips_layer_accum": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "saspa_conv3x3_small_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "saspa_conv3x3_small_bf16": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P]),
    "saspa_resize_area_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "saspa_resize_area_u8": (c_int, [_P, c_int, c_int, c_int, _P, c_int, c_int, _P, c_size_t, _P]),
    "saspa_resize_lanczos4_workspace_bytes": (c_size_t, [c_int, c_int]),
    "saspa_resize_lanczos4_u8": (c_int, [_P, c_int, c_int, c_int, _P, c_int, c_int, _P, c_size_t, _P]),
    "saspa_hed_fuse_u8": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
}

# kernel-selection overrides for tests / A-B timing (saspa_aug_b200/csrc/tuning_hooks.h): not part of the product ABI
TUNING_HOOKS = {name: (c_int, [c_int]) for name in ("saspa_attention_impl", "saspa_conv_impl", "saspa_groupnorm_impl", "saspa_gemm_force_ctas", "saspa_gemm_reverse_m",
                                                    "saspa_gemm_force_bn")}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise SaspaError(
                f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the hot path)"
            )
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in list(SIGNATURES.items()) + list(TUNING_HOOKS.items()):
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().saspa_last_error_string()
        raise SaspaError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
