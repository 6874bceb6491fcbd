"""ctypes binding of libmvae_b200.so (the C ABI declared in include/mvae_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a
``MvaeError`` is raised.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmvae_b200.so")

PREC_TF32 = 0
PREC_3XTF32 = 1
EPI_STORE, EPI_BIAS_SWISH, EPI_MUL_DSWISH = 0, 1, 2
GEMM_MAX_BATCH = 4
GEMM_MAX_CHAIN = 16
GEMM_CHAIN_WS_HEADER = 2


class MvaeError(RuntimeError):
    pass


class ConvView(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("lower_h", C.c_int32), ("lower_w", C.c_int32), ("upper_h", C.c_int32), ("upper_w", C.c_int32),
                ("stride", C.c_int32), ("taps_h", C.c_int32), ("taps_w", C.c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_mn_major", C.c_int32),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_mn_major", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("C", C.c_void_p), ("ldc", C.c_int64),
        ("bias", C.c_void_p),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("out2", C.c_void_p), ("ldout2", C.c_int64),
        ("colsum", C.c_void_p),
        ("epilogue", C.c_int32), ("split_k", C.c_int32), ("accumulate", C.c_int32),
        ("split_ws", C.c_void_p),
        ("B_lo", C.c_void_p),
        ("a_view", ConvView), ("b_view", ConvView),
        ("b_tap_slots", C.c_int32), ("b_tap_k", C.c_int32), ("b_tap_mn", C.c_int32),
        ("b_tap_table", C.c_int32 * 16),
        ("rowmap_IH", C.c_int32), ("rowmap_IW", C.c_int32), ("rowmap_s", C.c_int32), ("rowmap_py", C.c_int32),
        ("rowmap_px", C.c_int32),
    ]


_P, _I, _L, _F, _U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64

# name -> argtypes ; every function returns int status unless listed in _SPECIAL_RESTYPE
SIGNATURES = {
    "mvae_gemm_batch": [C.POINTER(GemmDesc), _I, _I, _P],
    "mvae_gemm_chain": [C.POINTER(GemmDesc), C.POINTER(C.c_int32), _I, _P, _L, _I, _P],
    "mvae_split_lo": [_P, _P, _L, _P],
    "mvae_linear_fwd": [_P, _L, _P, _L, _P, _P, _L, _P, _L, _I, _I, _I, _I, _P],
    "mvae_linear_dgrad": [_P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _I, _P],
    "mvae_linear_wgrad": [_P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _I, _P],
    "mvae_colsum_accumulate": [_P, _L, _P, _I, _I, _P],
    "mvae_swish_fwd": [_P, _P, _L, _P],
    "mvae_swish_bwd": [_P, _P, _P, _L, _P],
    "mvae_embedding_swish_fwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "mvae_embedding_swish_bwd": [_P, _P, _P, _L, _P, _I, _I, _I, _P],
    "mvae_poe_fwd": [C.POINTER(_P), C.POINTER(_P), _L, _I, C.POINTER(C.c_uint32), _I, _I, _I, _I, _I, _P, _P, _U64,
                     _U64, _P, _P, _L, _P, _P, _P, _P],
    "mvae_poe_bwd": [C.POINTER(_P), C.POINTER(_P), _L, _I, C.POINTER(C.c_uint32), _I, _I, _I, _I, _I, _P, _P, _L, _P,
                     _P, _F, _P, C.POINTER(_P), C.POINTER(_P), _L, _P],
    "mvae_poe_fwd_g": [C.POINTER(_P), C.POINTER(_P), _L, _I, C.POINTER(_P), C.POINTER(C.c_uint32), _I, _I, _I, _I, _I, _P, _P,
                       _U64, _U64, _P, _P, _L, _P, _P, _P, _P],
    "mvae_poe_bwd_g": [C.POINTER(_P), C.POINTER(_P), _L, _I, C.POINTER(_P), C.POINTER(C.c_uint32), _I, _I, _I, _I, _I, _P, _P,
                       _L, _P, _P, _F, _P, C.POINTER(_P), C.POINTER(_P), _L, _P],
    "mvae_label_table_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mvae_label_table_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "mvae_kl_fwd_bwd": [_P, _P, _P, _P, _L, _F, _P, _P],
    "mvae_reparam_fwd": [_P, _P, _P, _P, _U64, _U64, _P, _L, _P],
    "mvae_reparam_bwd": [_P, _P, _P, _P, _L, _P],
    "mvae_bce_logits_fwd_bwd": [_P, _L, _P, _L, _I, _P, _L, _I, _I, _F, _P, _I, _P, _L, _P],
    "mvae_ce_fwd_bwd": [_P, _L, _P, _I, _P, _L, _I, _I, _F, _P, _I, _P, _L, _P],
    "mvae_im2col_k4s2p1": [_P, _P, _L, _I, _I, _I, _I, _P],
    "mvae_col2im_k4s2p1": [_P, _L, _P, _P, _P, _I, _I, _I, _I, _P],
    "mvae_im2col_k4": [_P, _P, _L, _I, _I, _I, _I, _I, _I, _P],
    "mvae_col2im_k4": [_P, _L, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "mvae_conv_k4s2p1_cin_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mvae_conv_k4s2p1_cin_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mvae_convt_k4s2p1_cout_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mvae_convt_k4s2p1_cout_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "mvae_bn_stats": [_P, _L, _I, _I, _I, _P, _P],
    "mvae_bn_finalize": [_P, _I, _I, _I, _F, _F, _P, _P, _P, _P, C.POINTER(C.c_int32), _I, _P],
    "mvae_bn_eval_stats": [_P, _P, _I, _I, _F, _P, _P, _P],
    "mvae_bn_apply": [_P, _L, _P, _L, _I, _I, _I, _P, _P, _P, _P, _I, _P],
    "mvae_bn_bwd": [_P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P],
    "mvae_dropout_fwd": [_P, _I, _P, _P, _P, _I, _I, _F, _U64, _P, _P],
    "mvae_dropout_bwd": [_P, _P, _P, _I, _I, _I, _F, _P],
    "mvae_nchw_to_nhwc": [_P, _P, _I, _I, _I, _P],
    "mvae_gather_batch_u8": [_P, _L, _P, _P, _I, _P, _L, _P, _P],
    "mvae_adam_flat": [_P, _P, _P, _P, _L, _F, _P, _F, _F, _F, _F, _P, _P],
    "mvae_allreduce_adam_p2p": [C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _P, _P, _L, _I, _P, _I, _I, _F, _P, _F, _F, _F,
                                _P, _P],
    "mvae_elbo_finalize": [_P, _P, _P, _I, _F, _F, _F, _P, _F, _P, _P],
}
_SPECIAL = {
    "mvae_version": ([], C.c_int),
    "mvae_last_error": ([], C.c_char_p),
    "mvae_launch_count": ([], C.c_uint64),
    "mvae_device_sm_count": ([], C.c_int),
}
EXPORTED = sorted(list(SIGNATURES) + list(_SPECIAL))

_lib = None


def load() -> C.CDLL:
    """dlopen libmvae_b200.so; raises MvaeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MvaeError(
            f"{LIB_PATH} is missing: build it with `python -m multimodal_vae_public_b200.build` "
            "(there is no CPU/eager fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name, (argtypes, restype) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mvae_last_error().decode("utf-8", "replace")
        raise MvaeError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(load().mvae_launch_count())
