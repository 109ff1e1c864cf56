"""ctypes binding of libsww.so (the C ABI declared in include/sww.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present
when a context is created, the package raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsww.so")

u64, i32, f64 = C.c_uint64, C.c_int32, C.c_double
pd = C.POINTER(C.c_double)


class Geometry(C.Structure):
    _fields_ = [("grid", i32 * 3), ("box", f64 * 3), ("lmax", i32), ("n_k", i32),
                ("k_lo", pd), ("k_hi", pd), ("rax", pd * 3), ("kax", pd * 3), ("mas", pd * 3),
                ("device", i32)]


ctxp = C.c_void_p

# name -> argtypes ; every function returns int except sww_last_error
SIGNATURES = {
    "sww_abi_version": [],
    "sww_ctx_create": [C.POINTER(Geometry), C.POINTER(ctxp)],
    "sww_ctx_destroy": [ctxp],
    "sww_ctx_set_stream": [ctxp, u64],
    "sww_workspace_bytes": [ctxp, C.c_int, C.POINTER(u64)],
    "sww_bind_workspace": [ctxp, u64, u64],
    "sww_launch_counts": [ctxp, C.POINTER(u64), C.POINTER(u64)],
    "sww_set_fft_backend": [ctxp, C.c_int],
    "sww_row_mesh": [ctxp, C.POINTER(i32)],
    "sww_profile_enable": [ctxp, C.c_int],
    "sww_profile_report": [ctxp, C.c_char_p, u64],
    "sww_ylm_cache_bytes": [ctxp, C.POINTER(u64)],
    "sww_set_ylm_cache": [ctxp, u64],
    "sww_set_fields": [ctxp, u64, u64, u64, f64, f64],
    "sww_set_include_pix": [ctxp, C.c_int],
    "sww_mas_filter": [ctxp, u64, C.c_int, u64],
    "sww_fft_c2c": [ctxp, u64, u64, C.c_int],
    "sww_fft_r2c": [ctxp, u64, u64],
    "sww_fft_c2r": [ctxp, u64, u64],
    "sww_apply_cinv_fkp": [ctxp, u64, u64, f64, f64, f64, u64],
    "sww_apply_n": [ctxp, u64, u64, f64, f64, u64],
    "sww_ylm_project": [ctxp, u64, C.c_int, u64],
    "sww_apply_c_alpha_single": [ctxp, u64, u64, f64, C.c_int, C.c_int, C.c_int, u64],
    "sww_get_binmap": [ctxp, u64],
    "sww_paint": [ctxp, u64, u64, u64, pd, C.c_int, f64, u64],
    "sww_grf_legacy": [ctxp, C.c_uint32, u64, u64, u64, u64],
    "sww_grf_check": [ctxp],
    "sww_grf_legacy_batch": [ctxp, C.POINTER(C.c_uint32), C.c_int, u64, C.POINTER(u64), u64, u64],
    "sww_pk_fisher_realisation": [ctxp, u64, u64, u64],
    "sww_pk_data_q": [ctxp, u64, u64],
    "sww_mc_finalize": [ctxp, u64, u64, i32, i32],
    "sww_set_fiducial_pk": [ctxp, u64],
    "sww_ml_apply": [ctxp, C.c_int, u64, u64],
    "sww_ml_apply_cinv": [ctxp, u64, C.c_int, f64, f64, u64, C.POINTER(i32), C.POINTER(i32)],
    "sww_pk_q_from_cinv": [ctxp, u64, u64],
    "sww_pk_fisher_realisation_ml": [ctxp, u64, C.c_int, f64, u64, u64, C.POINTER(i32)],
    "sww_bk_set_triangles": [ctxp, C.POINTER(i32), i32],
    "sww_bk_gmaps": [ctxp, u64, C.c_int, u64, u64],
    "sww_bk_fisher_pair": [ctxp, u64, u64, u64, u64, u64, u64, u64, u64],
    "sww_bk_delta": [ctxp, u64],
    "sww_bk_q3": [ctxp, u64, u64],
    "sww_bk_q1_accumulate": [ctxp, u64, u64, u64, f64, u64],
    "sww_gram_f64": [ctxp, u64, u64, i32, i32, u64, C.c_int, u64],
}

WS_OPS, WS_PK_FISHER, WS_PK_DATA, WS_BK_FISHER, WS_BK_DATA, WS_PK_ML = range(6)

_lib = None


def load():
    """Load libsww.so and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Exception("libsww.so is not built (%s): run `python -m spectra_without_windows_b200.build`; "
                        "this package has no CPU path" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.sww_last_error.restype = C.c_char_p
    lib.sww_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the symbol is not exported
        fn.restype = C.c_int
        fn.argtypes = args
    if lib.sww_abi_version() != 1:
        raise Exception("libsww.so ABI mismatch")
    _lib = lib
    return lib


def check(rc):
    """Turn a non-zero return code into the bare Exception the reference scripts use."""
    if rc != 0:
        raise Exception("libsww: " + _lib.sww_last_error().decode())
