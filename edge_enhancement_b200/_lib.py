"""ctypes binding of libedge_b200.so (include/edge_b200.h).  No torch types cross this boundary.

The product has no CPU fallback: if the shared library is missing and cannot be built, importing
this module raises; if a call fails, `check` raises RuntimeError with ee_last_error().
"""
import ctypes
import os
import threading

from . import _build

EE_VARIANT_STEP125, EE_VARIANT_CANNY, EE_VARIANT_BPDA = 0, 1, 2
EE_LAYOUT_NCHW = 0
EE_FLAG_NAN_COMPAT = 1


class EEParams(ctypes.Structure):
    """Mirror of `struct EEParams` in include/edge_b200.h."""
    _fields_ = [("variant", ctypes.c_int32), ("layout", ctypes.c_int32),
                ("gauss", ctypes.c_float * 9), ("sobel", ctypes.c_float * 9),
                ("alpha", ctypes.c_float), ("low_thr", ctypes.c_float), ("high_thr", ctypes.c_float),
                ("has_low", ctypes.c_int32), ("has_high", ctypes.c_int32),
                ("hysteresis", ctypes.c_int32), ("flags", ctypes.c_int32)]


class EEStrides(ctypes.Structure):
    """Mirror of `struct EEStrides`: element strides (n, c, h, w) = torch.Tensor.stride() of a [B,C,H,W] tensor."""
    _fields_ = [("n", ctypes.c_int64), ("c", ctypes.c_int64), ("h", ctypes.c_int64), ("w", ctypes.c_int64)]


_vp, _i, _i64, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
_pp = ctypes.POINTER(EEParams)
_sp = ctypes.POINTER(EEStrides)

# name -> argtypes ; every function returns int unless listed in _RESTYPE
SIGNATURES = {
    "ee_edge_fwd_f32": [_vp, _vp, _i, _i, _i, _i, _pp, _vp],
    "ee_edge_bwd_f32": [_vp, _vp, _vp, _i, _i, _i, _i, _pp, _vp],
    "ee_edge_blend_fwd_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _pp, _f, _vp],
    "ee_edge_blend_bwd_f32": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _pp, _f, _vp],
    "ee_edge_fwd_strided_f32": [_vp, _sp, _vp, _sp, _i, _i, _i, _i, _pp, _vp],
    "ee_edge_bwd_strided_f32": [_vp, _sp, _vp, _sp, _vp, _sp, _i, _i, _i, _i, _pp, _vp],
    "ee_edge_blend_fwd_strided_f32": [_vp, _sp, _vp, _sp, _vp, _sp, _vp, _sp, _i, _i, _i, _i, _pp, _f, _vp],
    "ee_edge_blend_bwd_strided_f32": [_vp, _sp, _vp, _sp, _vp, _sp, _vp, _sp, _vp, _sp, _i, _i, _i, _i, _pp, _f, _vp],
    "ee_aux_bytes": [_i, _i, _i, _i, _i],
    "ee_pgd_linf_step_f32": [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _vp],
    "ee_fgsm_step_f32": [_vp, _vp, _vp, _i64, _f, _f, _f, _vp],
    "ee_free_at_step_f32": [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _vp],
    "ee_cw_linf_step_f32": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _vp],
    "ee_pgd_l2_step_f32": [_vp, _vp, _vp, _vp, _i, _i64, _f, _f, _vp],
    "ee_gf_blend_fwd_f32": [_vp, _vp, _vp, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_float), _f, _vp],
    "ee_gf_blend_bwd_f32": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_float), _f, _vp],
    "ee_edge_pgd_iteration_f32": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _pp, _f, _f, _f, _vp],
    "ee_add_clamp_f32": [_vp, _vp, _vp, _i64, _f, _f, _vp],
    "ee_avmixup_mix_f32": [_vp, _vp, _vp, _vp, _i, _i64, _f, _vp],
    "ee_to_compare_fwd_f32": [_vp, _vp, _i64, _f, _vp],
    "ee_to_compare_bwd_f32": [_vp, _vp, _vp, _i64, _f, _vp],
    "ee_to_eq_fwd_f32": [_vp, _vp, _i64, _vp],
    "ee_to_eq_bwd_f32": [_vp, _vp, _vp, _i64, _vp],
    "ee_safe_sign_fwd_f32": [_vp, _vp, _i64, _vp],
    "ee_safe_sign_bwd_f32": [_vp, _vp, _vp, _i64, _vp],
    "ee_add_square_fwd_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp],
    "ee_add_square_bwd_f32": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp],
    "ee_hfs_f32": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _f, _vp],
    "ee_hfs_supported": [_i, _i],
    "ee_hfs_tc_f32": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _f, _vp],
    "ee_hfs_tc_supported": [_i, _i],
    "ee_last_error": [],
    "ee_version": [],
    "ee_set_tuning": [_i, _i, _i],
}
_RESTYPE = {"ee_last_error": ctypes.c_char_p, "ee_aux_bytes": ctypes.c_size_t}

_LIB = None
_LOCK = threading.Lock()


def lib_path():
    return _build.LIB


def load():
    """Load (building first if the .so is missing) and type every exported symbol.  Thread-safe: DataParallel
    worker threads and autograd engine threads may race for the first call."""
    global _LIB
    if _LIB is not None:
        return _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        # EDGE_B200_LIB is a TEST-ONLY override (A/B timing of two kernel builds): it loads whatever .so it names
        path = os.environ.get("EDGE_B200_LIB") or _build.LIB
        if not os.path.exists(path):
            path = _build.build()            # raises if nvcc is unavailable: no silent fallback
        L = ctypes.CDLL(path)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)            # AttributeError if the library lacks a declared symbol
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, ctypes.c_int)
        _LIB = L
    return L


def check(rc, what):
    if rc != 0:
        msg = load().ee_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))
