"""ctypes binding of oracle/libsto_oracle.so (the CPU restatement).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, and
from nowhere else: the product never routes through it.  `build()` compiles the C file with gcc.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsto_oracle.so")
MAX_BREAKS = 32
_dp = C.POINTER(C.c_double)


class OracleVehicle(C.Structure):
    _fields_ = [("max_lon_acc", C.c_double), ("max_lon_dcc", C.c_double), ("max_left_acc", C.c_double),
                ("max_right_acc", C.c_double), ("max_speed", C.c_double), ("max_jerk", C.c_double),
                ("n_acc", C.c_int32), ("n_dcc", C.c_int32),
                ("acc_x", C.c_double * MAX_BREAKS), ("acc_c", (C.c_double * (MAX_BREAKS - 1)) * 4),
                ("dcc_x", C.c_double * MAX_BREAKS), ("dcc_c", (C.c_double * (MAX_BREAKS - 1)) * 4)]


def make_vehicle(scalars, acc_x, acc_c, dcc_x, dcc_c):
    """scalars = (max_lon_acc, max_lon_dcc, max_left, max_right, vmax, jerk); *_x, *_c = PPoly.x / PPoly.c."""
    v = OracleVehicle()
    (v.max_lon_acc, v.max_lon_dcc, v.max_left_acc, v.max_right_acc, v.max_speed, v.max_jerk) = [float(s) for s in scalars]
    for name, x, c in (("acc", acc_x, acc_c), ("dcc", dcc_x, dcc_c)):
        x = np.asarray(x, dtype=np.float64)
        c = np.asarray(c, dtype=np.float64)
        assert c.shape == (4, len(x) - 1) and len(x) <= MAX_BREAKS
        setattr(v, "n_" + name, len(x))
        xs = getattr(v, name + "_x")
        cs = getattr(v, name + "_c")
        for i, xv in enumerate(x):
            xs[i] = xv
        for k in range(4):
            for i in range(len(x) - 1):
                cs[k][i] = c[k, i]
    return v


def build(force=False):
    src = os.path.join(HERE, "sto_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B", "libsto_oracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.sto_oracle_ppoly.restype = C.c_double
        L.sto_oracle_ppoly.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_double]
        L.sto_oracle_maxlat.restype = C.c_double
        L.sto_oracle_maxlat.argtypes = [C.POINTER(OracleVehicle), C.c_double, C.c_int]
        L.sto_oracle_fit_periodic_cubic.argtypes = [_dp, _dp, C.c_int, _dp, _dp, _dp]
        L.sto_oracle_bspline_eval.restype = None
        L.sto_oracle_bspline_eval.argtypes = [_dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, _dp]
        L.sto_oracle_sample.argtypes = [_dp, C.c_int, _dp, _dp, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_int]
        L.sto_oracle_qss.argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.POINTER(OracleVehicle), _dp, _dp, _dp, _dp,
                                     _dp, _dp, C.POINTER(C.c_int64), C.c_int]
        L.sto_oracle_lap_from_offsets.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, C.c_int,
                                                  C.POINTER(OracleVehicle), _dp, C.c_int]
        L.sto_oracle_lap_batch.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int, _dp, _dp, C.c_int,
                                           C.POINTER(OracleVehicle), _dp, C.POINTER(C.c_int32), C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ppoly(x, c, v):
    x, c = _f64(x), _f64(c)
    return lib().sto_oracle_ppoly(_p(x), _p(c), len(x), c.shape[0], float(v))


def maxlat(veh, lon, ref_pow=1):
    return lib().sto_oracle_maxlat(C.byref(veh), float(lon), int(ref_pow))


def set_fit_solver(name):
    """'fitpack' (default): FITPACK's fpclos Givens sweep, bit-identical to scipy splprep(per=1, s=0) = the reference;
    'blocks': 32-block elimination + cyclic reduction; 'thomas': one-lane Thomas + Sherman-Morrison.  The CUDA library
    has the same three (sto_set_fit_solver).  Applies to fit_periodic_cubic and lap_batch."""
    lib().sto_oracle_set_fit_solver({"fitpack": 2, "blocks": 1, "thomas": 0}[name])


def fit_periodic_cubic(points):
    pts = _f64(points)
    M = pts.shape[0]
    px, py = _f64(pts[:, 0]), _f64(pts[:, 1])
    t, cx, cy = np.empty(M + 7), np.empty(M + 3), np.empty(M + 3)
    rc = lib().sto_oracle_fit_periodic_cubic(_p(px), _p(py), M, _p(t), _p(cx), _p(cy))
    if rc != 0:
        raise RuntimeError(f"oracle fit failed rc={rc}")
    return t, cx, cy


def bspline_eval(t, c, k, xs, der=0):
    t, c, xs = _f64(t), _f64(c), _f64(xs)
    out = np.empty(len(xs))
    lib().sto_oracle_bspline_eval(_p(t), len(t), _p(c), int(k), _p(xs), len(xs), int(der), _p(out))
    return out


def sample(t, cx, cy, k, ts, ref_pow=1):
    t, cx, cy, ts = _f64(t), _f64(cx), _f64(cy), _f64(ts)
    N = len(ts)
    X, Y, YAW, R = (np.empty(N) for _ in range(4))
    rc = lib().sto_oracle_sample(_p(t), len(t), _p(cx), _p(cy), int(k), _p(ts), N, _p(X), _p(Y), _p(YAW), _p(R),
                                 int(ref_pow))
    if rc != 0:
        raise RuntimeError(f"oracle sample failed rc={rc}")
    return X, Y, YAW, R


def qss(X, Y, R, sinb, veh, ref_pow=1):
    X, Y, R, sinb = _f64(X), _f64(Y), _f64(R), _f64(sinb)
    N = len(X)
    v, a, lat, flag, tseg = (np.empty(N) for _ in range(5))
    lap = C.c_double()
    stats = (C.c_int64 * 8)()
    rc = lib().sto_oracle_qss(_p(X), _p(Y), _p(R), _p(sinb), N, C.byref(veh), _p(v), _p(a), _p(lat), _p(flag),
                              _p(tseg), C.byref(lap), stats, int(ref_pow))
    return dict(status=rc, v=v, a=a, lat=lat, flag=flag, time=tseg, lap=lap.value,
                steps=stats[0], iters=stats[1], spawned=stats[2], peak_rows=stats[3], effective=stats[4])


def qss_sweep(X, Y, R, sinb, veh, rounds=2, ref_pow=0):
    """FAST-MODE checker (not the reference's algorithm): `rounds` x (backward lap, forward lap) with the reference's
    step operator, the CPU twin of the library's sto_lap_time_fast_f64."""
    X, Y, R, sinb = _f64(X), _f64(Y), _f64(R), _f64(sinb)
    N = len(X)
    v, a, lat, tseg = (np.empty(N) for _ in range(4))
    lap = C.c_double()
    L = lib()
    L.sto_oracle_qss_sweep.restype = C.c_int
    L.sto_oracle_qss_sweep.argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.POINTER(OracleVehicle), C.c_int, _dp, _dp, _dp, _dp,
                                       C.POINTER(C.c_double), C.c_int]
    rc = L.sto_oracle_qss_sweep(_p(X), _p(Y), _p(R), _p(sinb), N, C.byref(veh), int(rounds), _p(v), _p(a), _p(lat),
                                _p(tseg), C.byref(lap), int(ref_pow))
    return dict(status=rc, v=v, a=a, lat=lat, time=tseg, lap=lap.value)


def lap_batch(centre_x, centre_y, normal_x, normal_y, offsets, ts, sinb, veh, n_threads=1, ref_pow=1):
    """offsets[B, M] (candidate-major, host).  Returns lap[B], status[B]."""
    cx, cy, nx, ny = _f64(centre_x), _f64(centre_y), _f64(normal_x), _f64(normal_y)
    off, ts, sinb = _f64(offsets), _f64(ts), _f64(sinb)
    B, M = off.shape
    lap = np.empty(B)
    status = np.empty(B, dtype=np.int32)
    rc = lib().sto_oracle_lap_batch(_p(cx), _p(cy), _p(nx), _p(ny), _p(off), M, B, _p(ts), _p(sinb), len(ts),
                                    C.byref(veh), _p(lap), status.ctypes.data_as(C.POINTER(C.c_int32)),
                                    int(n_threads), int(ref_pow))
    if rc != 0:
        raise RuntimeError(f"oracle batch failed rc={rc}")
    return lap, status
