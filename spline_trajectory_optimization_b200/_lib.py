"""ctypes binding of libsto_b200.so (include/sto_b200.h).  No CPU fallback: if the CUDA library is missing
or fails to load, every entry point raises."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libsto_b200.so")
MAX_BREAKS = 32
QSS_PLAIN, QSS_MEMO = 0, 1

CAND_ZERO_SPEED, CAND_NAN, CAND_ROW_OVERFLOW, CAND_NO_CONVERGENCE, CAND_DEGENERATE_FIT = 1, 2, 4, 8, 16

_vp = C.c_void_p


class StoVehicle(C.Structure):
    """sto_vehicle_f64"""
    _fields_ = [("max_lon_acc", C.c_double), ("max_lon_dcc", C.c_double), ("max_left_acc", C.c_double),
                ("max_right_acc", C.c_double), ("max_speed", C.c_double), ("max_jerk", C.c_double),
                ("n_acc", C.c_int32), ("n_dcc", C.c_int32),
                ("acc_x", C.c_double * MAX_BREAKS), ("acc_c", (C.c_double * (MAX_BREAKS - 1)) * 4),
                ("dcc_x", C.c_double * MAX_BREAKS), ("dcc_c", (C.c_double * (MAX_BREAKS - 1)) * 4)]


class StoProfileOut(C.Structure):
    """sto_profile_out_f64"""
    _fields_ = [("speed", _vp), ("lon_acc", _vp), ("lat_acc", _vp), ("time", _vp), ("owner", _vp)]


class StoFastOut(C.Structure):
    """sto_fast_out_f64"""
    _fields_ = [(k, _vp) for k in ("cx", "cy", "x", "y", "yaw", "radius", "speed", "lon_acc", "lat_acc", "time")]


class StoError(RuntimeError):
    pass


_SIGNATURES = {
    "sto_abi_version": (C.c_int, []),
    "sto_last_error": (C.c_char_p, []),
    "sto_last_qss_kernel": (C.c_char_p, []),
    "sto_device_count": (C.c_int, []),
    "sto_release": (C.c_int, []),
    "sto_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
    "sto_fit_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "sto_fit_periodic_cubic_f64": (C.c_int, [_vp] * 7 + [C.c_int] * 3 + [_vp] * 4 + [_vp, C.c_size_t, _vp]),
    "sto_sample_f64": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int] + [_vp] * 6 + [_vp]),
    "sto_sample_spline_f64": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int] + [_vp] * 4 + [_vp]),
    "sto_fill_bounds_f64": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_double, _vp, _vp, _vp, _vp]),
    "sto_arc_sections_f64": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "sto_sample_splines_f64": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int] + [_vp] * 6 + [_vp]),
    "sto_lap_splines_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "sto_lap_time_splines_f64": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(StoVehicle), C.c_int, _vp, _vp, _vp, C.c_size_t, _vp]),
    "sto_qss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "sto_qss_f64": (C.c_int, [_vp] * 4 + [C.c_int] * 3 + [C.POINTER(StoVehicle), C.c_int, _vp, _vp,
                              C.POINTER(StoProfileOut), _vp, _vp, C.c_size_t, _vp]),
    "sto_lap_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "sto_lap_time_f64": (C.c_int, [_vp] * 7 + [C.c_int] * 4 + [C.POINTER(StoVehicle), C.c_int, _vp, _vp, _vp,
                                   C.c_size_t, _vp]),
    "sto_lap_time_host_f64": (C.c_int, [_vp] * 7 + [C.c_int] * 3 + [C.POINTER(StoVehicle), C.c_int, _vp, _vp,
                                        C.c_int, C.c_size_t]),
    "sto_fit_lsq_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "sto_fit_periodic_lsq_f64": (C.c_int, [_vp] * 7 + [C.c_int] * 3 + [_vp, C.c_int, C.c_int] + [_vp] * 5 +
                                 [C.c_size_t, _vp]),
    "sto_fast_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "sto_lap_time_fast_f64": (C.c_int, [_vp] * 7 + [C.c_int] * 4 + [C.POINTER(StoVehicle), C.c_int, _vp, _vp, _vp, _vp,
                                        C.c_size_t, C.c_int, _vp]),
    "sto_set_stage_timing": (C.c_int, [C.c_int]),
    "sto_set_fit_solver": (C.c_int, [C.c_int]),
    "sto_get_fit_solver": (C.c_int, []),
    "sto_fit_solver_lanes": (C.c_int, [C.c_int, C.c_int]),
    "sto_last_stage_ms": (C.c_int, [C.POINTER(C.c_float)]),
    "sto_selftest_fp64": (C.c_int, [C.c_ulonglong] + [C.POINTER(C.c_longlong)] * 3),
    "sto_measure_fp64_peak": (C.c_int, [C.POINTER(C.c_double)]),
    "sto_argmin_f64": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp]),
    "sto_argmin_pair_f64": (C.c_int, [_vp, _vp, C.c_int, C.c_int64, _vp, _vp, _vp]),
    "sto_argmin_pairs_f64": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "sto_transpose_f64": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)
FIT_THOMAS, FIT_BLOCKS, FIT_FITPACK = 0, 1, 2   # sto_set_fit_solver

_lib = None


def load():
    """Load libsto_b200.so; raises StoError when it has not been built (python -m ...build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StoError(f"{LIB_PATH} not built: run `python -m spline_trajectory_optimization_b200.build` "
                           "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.sto_abi_version() != 1:
            raise StoError("libsto_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise StoError(f"libsto_b200 error {rc}: {load().sto_last_error().decode()}")


def make_vehicle(scalars, acc_x, acc_c, dcc_x, dcc_c):
    """scalars = (max_lon_acc, max_lon_dcc, max_left, max_right, max_speed, max_jerk); tables = PPoly x / c."""
    v = StoVehicle()
    (v.max_lon_acc, v.max_lon_dcc, v.max_left_acc, v.max_right_acc, v.max_speed, v.max_jerk) = [float(s) for s in scalars]
    for name, x, c in (("acc", acc_x, acc_c), ("dcc", dcc_x, dcc_c)):
        x = np.asarray(x, dtype=np.float64)
        c = np.asarray(c, dtype=np.float64)
        if x.ndim != 1 or not (2 <= len(x) <= MAX_BREAKS):
            raise ValueError(f"{name} table needs 2..{MAX_BREAKS} break points")
        if c.shape[1] != len(x) - 1 or c.shape[0] > 4:
            raise ValueError(f"{name} PPoly coefficients must be [<=4][{len(x) - 1}]")
        if c.shape[0] < 4:  # lower-order PPoly: pad the high-order rows with zeros
            c = np.vstack([np.zeros((4 - c.shape[0], c.shape[1])), c])
        setattr(v, "n_" + name, len(x))
        for i, xv in enumerate(x):
            getattr(v, name + "_x")[i] = xv
        for k in range(4):
            for i in range(len(x) - 1):
                getattr(v, name + "_c")[k][i] = c[k, i]
    return v
