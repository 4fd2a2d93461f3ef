"""ctypes binding of libunet_b200.so (the C-ABI declared in include/unet_b200.h).

The reference has no FFI: TensorFlow executes its Keras graph.  This module is the only place Python touches
native code; every call goes to a hand-written sm_100a kernel.  A missing library is a hard error (no fallback).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libunet_b200.so"

UNET_F32, UNET_BF16 = 0, 1
EPI_NONE, EPI_AFFINE, EPI_AFFINE_RELU, EPI_STATS, EPI_CONVT, EPI_HEAD = 0, 1, 2, 3, 4, 5


class UnetError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.code = code


class Dropout(C.Structure):
    _fields_ = [("rate", C.c_float), ("seed", C.c_uint32), ("ctot", C.c_int64), ("c0", C.c_int64),
                ("seed_dev", C.c_void_p)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64),
        ("C", C.c_void_p), ("ldc", C.c_int64),
        ("a_trans", C.c_int), ("b_trans", C.c_int),
        ("in_dtype", C.c_int), ("out_dtype", C.c_int),
        ("accumulate", C.c_int), ("epilogue", C.c_int),
        ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("colsum", C.c_void_p), ("colsq", C.c_void_p),
        ("convt_H", C.c_int), ("convt_W", C.c_int),
        ("drop", Dropout),
        ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_out", C.c_void_p), ("head_classes", C.c_int),
        ("A2", C.c_void_p), ("lda2", C.c_int64), ("k_split", C.c_int64),
        ("B2", C.c_void_p), ("ldb2", C.c_int64), ("n_split", C.c_int64),
        ("A_lo", C.c_void_p), ("lda_lo", C.c_int64), ("B_lo", C.c_void_p), ("ldb_lo", C.c_int64),
    ]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_dp = C.POINTER(Dropout)

# name -> argtypes (all return int unless listed in _RESTYPES)
_SIGNATURES = {
    "unet_version": [],
    "unet_sm_arch": [],
    "unet_last_error": [],
    "unet_device_check": [_i],
    "unet_dwconv3x3_fwd": [_vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp, _vp, _dp, _vp, _vp],
    "unet_dwconv3x3_bwd_weight": [_vp, _i64, _vp, _i64, _vp, _i, _i, _i, _i, _i, _vp],
    "unet_bn_bwd_coef": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "unet_bn_bwd_wgrad_combine": [_vp, _vp, _vp, _vp, _i, _i, _vp],
    "unet_pw_bwd_fused": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _vp],
    "unet_dwconv3x3_bwd": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _i, _vp, _dp, _i, _vp, _vp, _vp, _i, _vp, _vp],
    "unet_stem_fwd": [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp],
    "unet_stem_bwd_folded": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp],
    "unet_stem_bwd": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "unet_sepconv_fused_fwd": [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _i, _vp, _i64, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i64, _vp],
    "unet_gemm_simt": [C.POINTER(GemmArgs), _vp],
    "unet_gemm_tc": [C.POINTER(GemmArgs), _vp],
    "unet_bn_fold": [_vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _vp],
    "unet_bn_finalize": [_vp, _vp, _i64, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "unet_bn_act": [_vp, _vp, _vp, _i, _vp, _i64, _vp, _i, _i, _i, _i, _i, _dp, _vp],
    "unet_bn_bwd_reduce": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _dp, _vp],
    "unet_bn_bwd_apply": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _dp, _vp],
    "unet_maxpool2x2_fwd": [_vp, _i64, _vp, _i, _i, _i, _i, _i, _vp],
    "unet_maxpool2x2_bwd": [_vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _vp, _dp, _vp],
    "unet_convt_bwd_gather": [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _dp, _vp],
    "unet_head_fwd": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _vp, _vp, _vp],
    "unet_seg_loss_finalize": [_vp, _i, _f, _i, _f, _vp, _vp, _vp],
    "unet_head_bwd": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _i, _i, _i, _vp, _vp, _vp, _vp],
    "unet_seg_sums": [_vp, _vp, _vp, _i64, _i64, _i, _vp],
    "unet_confusion_matrix_update": [_vp, _vp, _i64, _i, _vp, _vp],
    "unet_confusion_matrix_update_thr": [_vp, _vp, _f, _i64, _vp, _vp],
    "unet_sample_confusion_thr": [_vp, _vp, _f, _i64, _i64, _vp, _vp],
    "unet_split_tf32": [_vp, _i64, _i64, _i64, _vp, _vp, _i, _vp],
    "unet_adamw_step": [_vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "unet_step_advance": [_vp, _vp, _vp],
    "unet_cast_transpose_bf16": [_vp, _vp, _vp, _i, _i, _vp, _vp],
    "unet_cast_transpose_bf16_batched": [_vp, _vp, _i, _i64, _vp],
    "unet_cast": [_vp, _i, _vp, _i, _i64, _vp],
    "unet_preprocess_u8": [_vp, _i, _i, _i, _i64, _vp, _i, _i, _f, _vp],
    "unet_postprocess_mask": [_vp, _i, _i, _i64, _vp, _i, _i, _f, _vp],
    "unet_host_dropout_hash": [C.c_uint64, C.c_uint32],
}
_RESTYPES = {"unet_last_error": C.c_char_p, "unet_host_dropout_hash": C.c_uint32}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    """Load the shared library; raises if it has not been built (python __graft_entry__.py / build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "unet_b200 has no CPU or library fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(fn_name: str, rc: int) -> None:
    if rc != 0:
        msg = load().unet_last_error()
        raise UnetError(fn_name, rc, msg.decode() if msg else "")


def call(fn_name: str, *args) -> None:
    check(fn_name, getattr(load(), fn_name)(*args))
