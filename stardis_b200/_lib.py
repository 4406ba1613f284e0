"""ctypes binding of libstardis_b200.so (the C ABI in include/stardis_b200.h).

There is NO fallback: if the shared library is missing or no CUDA device is usable, importing callers get a
loud error.  (``python -m stardis_b200.build`` / ``__graft_entry__.build()`` produce the library.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libstardis_b200.so")

SD_OK = 0
LINEAR_STARK, QUADRATIC_STARK, VAN_DER_WAALS, RADIATION, VALD = 1, 2, 4, 8, 16
BUF_GAMMAS, BUF_DOPPLER, BUF_ALPHA_LINE, BUF_ALPHA_MOLECULE, BUF_TOTAL, BUF_F_NU, BUF_I_NUS = 1, 2, 3, 4, 5, 6, 7
BUF_LINE_STRENGTH = 8
BUF_NUS = 9
BUF_SOURCE0 = 16
SRC_BF, SRC_FF, SRC_RAYLEIGH, SRC_ELECTRON, SRC_TABLE0 = 0, 1, 2, 3, 4
MAX_TABLES = 8

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


class SdLines(C.Structure):
    _fields_ = [("n_lines", C.c_int64), ("nu", C.c_void_p), ("atomic_number", C.c_void_p), ("ion_number", C.c_void_p),
                ("ionization_energy", C.c_void_p), ("level_energy_upper", C.c_void_p), ("level_energy_lower", C.c_void_p),
                ("A_ul", C.c_void_p), ("mass", C.c_void_p), ("stark", C.c_void_p), ("waals", C.c_void_p),
                ("alpha_line", C.c_void_p)]


class SdTable(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("x", C.c_void_p), ("y", C.c_void_p),
                ("values", C.c_void_p), ("diag", C.c_void_p), ("depth_y", C.c_void_p), ("depth_scale", C.c_void_p)]


class SdContinuum(C.Structure):
    _fields_ = [("n_bf_levels", C.c_int32), ("bf_nu_cut", C.c_void_p), ("bf_prefix", C.c_void_p), ("ff_coef", C.c_void_p),
                ("ray_c4", C.c_void_p), ("ray_c6", C.c_void_p), ("ray_c8", C.c_void_p), ("electron", C.c_void_p),
                ("n_tables", C.c_int32), ("tables", SdTable * MAX_TABLES)]


# every symbol declared in include/stardis_b200.h: name -> (restype, argtypes)
_V = C.c_void_p
SIGNATURES = {
    "sd_create": (C.c_int, [C.POINTER(_V), C.c_int]),
    "sd_destroy": (None, [_V]),
    "sd_last_error": (C.c_char_p, [_V]),
    "sd_version": (C.c_char_p, []),
    "sd_set_stream": (C.c_int, [_V, _V]),
    "sd_synchronize": (C.c_int, [_V]),
    "sd_host_alloc": (C.c_int, [C.POINTER(_V), C.c_int64]),
    "sd_host_free": (C.c_int, [_V]),
    "sd_set_atmosphere": (C.c_int, [_V, C.c_int32, _V, _V, _V, C.c_double]),
    "sd_set_grid": (C.c_int, [_V, C.c_int64, _V, C.c_int64, C.c_int64]),
    "sd_set_lines": (C.c_int, [_V, C.POINTER(SdLines)]),
    "sd_calc_broadening": (C.c_int, [_V, C.c_uint32]),
    "sd_set_broadening": (C.c_int, [_V, _V, C.c_int32, _V]),
    "sd_calc_alpha_line": (C.c_int, [_V, C.c_int32]),
    "sd_calc_alpha_line_vald": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V, _V]),
    "sd_calc_alpha_line_levels": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V, _V, _V]),
    "sd_set_farfield": (C.c_int, [_V, C.c_int32]),
    "sd_set_line_stats": (C.c_int, [_V, C.c_int32]),
    "sd_line_stats": (C.c_int, [_V, _ip]),
    "sd_line_stats_ex": (C.c_int, [_V, _ip]),
    "sd_phase_times": (C.c_int, [_V, C.POINTER(C.c_float)]),
    "sd_calc_continuum": (C.c_int, [_V, C.POINTER(SdContinuum), C.c_uint32]),
    "sd_raytrace": (C.c_int, [_V, C.c_int32, _V, _V, C.c_int32, C.c_double, C.c_int32]),
    "sd_get": (C.c_int, [_V, C.c_int32, _V, C.c_int64]),
    "sd_get_row": (C.c_int, [_V, C.c_int32, C.c_int32, _V, C.c_int64]),
    "sd_set_total": (C.c_int, [_V, _V, C.c_int64]),
    "sd_buffer": (C.c_int, [_V, C.c_int32, C.POINTER(_V), C.POINTER(C.c_int64)]),
    "sd_ew_faddeeva": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V]),
    "sd_ew_voigt_profile": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V]),
    "sd_ew_doppler_width": (C.c_int, [_V, C.c_int64, _V, _V, _V, C.c_double, _V]),
    "sd_ew_n_effective": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V]),
    "sd_ew_gamma_linear_stark": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V]),
    "sd_ew_gamma_quadratic_stark": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V, _V, _V]),
    "sd_ew_gamma_van_der_waals": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V, _V, _V]),
    "sd_ew_blackbody": (C.c_int, [_V, C.c_int32, C.c_int64, _V, _V, _V]),
    "sd_ew_calc_weights": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V]),
    "sd_convolve1d_reflect": (C.c_int, [_V, C.c_int64, _V, C.c_int32, _V, _V]),
    "sd_bench_dfma": (C.c_int, [_V, C.c_int32, _dp]),
    "sd_bench_fp64": (C.c_int, [_V, C.c_int32, C.c_int32, _dp]),
    "sd_bench_fareval": (C.c_int, [_V, C.c_int32, C.c_int32, _dp]),
    "sd_debug_rcp": (C.c_int, [_V, C.c_int64, _V, _V, _V, _V]),
    "sd_launch_count": (C.c_int64, [_V]),
    "sd_timer_start": (C.c_int, [_V]),
    "sd_timer_stop": (C.c_int, [_V, C.POINTER(C.c_float)]),
}

_lib = None


class StardisB200Error(RuntimeError):
    pass


def load():
    """dlopen libstardis_b200.so and bind every entry point; raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StardisB200Error(
                f"{LIB_PATH} not found: build it with `python -m stardis_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(a):
    """Raw address of a numpy array / torch tensor / int / None, as c_void_p."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return int(a)


def f64(a):
    """C-contiguous float64 view/copy of a host array (torch tensors are passed through untouched)."""
    if hasattr(a, "data_ptr"):
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    if hasattr(a, "data_ptr"):
        return a
    return np.ascontiguousarray(a, dtype=np.int64)
