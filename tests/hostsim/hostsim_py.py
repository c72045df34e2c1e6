"""ctypes binding of tests/hostsim/libhostsim.so: the device functions compiled for the host.  TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhostsim.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
MAX_BREAKS = 32


class Vehicle(C.Structure):  # == sto_vehicle_f64 (include/sto_b200.h)
    _fields_ = [("max_lon_acc", C.c_double), ("max_lon_dcc", C.c_double), ("max_left_acc", C.c_double),
                ("max_right_acc", C.c_double), ("max_speed", C.c_double), ("max_jerk", C.c_double),
                ("n_acc", C.c_int32), ("n_dcc", C.c_int32),
                ("acc_x", C.c_double * MAX_BREAKS), ("acc_c", (C.c_double * (MAX_BREAKS - 1)) * 4),
                ("dcc_x", C.c_double * MAX_BREAKS), ("dcc_c", (C.c_double * (MAX_BREAKS - 1)) * 4)]


def make_vehicle(scalars, acc_x, acc_c, dcc_x, dcc_c):
    v = Vehicle()
    (v.max_lon_acc, v.max_lon_dcc, v.max_left_acc, v.max_right_acc, v.max_speed, v.max_jerk) = [float(s) for s in scalars]
    for name, x, c in (("acc", acc_x, acc_c), ("dcc", dcc_x, dcc_c)):
        x = np.asarray(x, dtype=np.float64); c = np.asarray(c, dtype=np.float64)
        setattr(v, "n_" + name, len(x))
        for i, xv in enumerate(x):
            getattr(v, name + "_x")[i] = xv
        for k in range(4):
            for i in range(len(x) - 1):
                getattr(v, name + "_c")[k][i] = c[k, i]
    return v


def build():
    srcs = [os.path.join(HERE, "hostsim.cpp"), os.path.join(HERE, "memo_instrument.h"), os.path.join(HERE, "memo3_proto.h")] + [
        os.path.join(HERE, "..", "..", "spline_trajectory_optimization_b200", "csrc", f)
        for f in ("sto_common.cuh", "sto_fast.cuh", "sto_fit.cuh", "sto_fit_lsq.cuh", "sto_eval.cuh", "sto_qss.cuh", "sto_qss_memo.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-std=c++17", "-o", LIB, srcs[0]])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def sm(a):
    """[B, n] candidate-major -> [n, B] sample-major contiguous."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)


def fit_points(points, split=1):
    """points[B, M, 2] -> u[B, M+1], cx[B, M+3], cy[B, M+3], status[B]"""
    pts = np.asarray(points, dtype=np.float64)
    B, M, _ = pts.shape
    px, py = sm(pts[:, :, 0]), sm(pts[:, :, 1])
    u, cx, cy = np.empty((M + 1, B)), np.empty((M + 3, B)), np.empty((M + 3, B))
    st = np.zeros(B, dtype=np.int32)
    lib().hostsim_fit(None, None, None, None, None, _p(px), _p(py), M, B, B, _p(u), _p(cx), _p(cy),
                      st.ctypes.data_as(_ip), int(split))
    return u.T.copy(), cx.T.copy(), cy.T.copy(), st


def fit_offsets(cenx, ceny, nrmx, nrmy, offsets, split=1):
    off = np.asarray(offsets, dtype=np.float64)
    B, M = off.shape
    o = sm(off)
    cenx, ceny, nrmx, nrmy = (np.ascontiguousarray(a, dtype=np.float64) for a in (cenx, ceny, nrmx, nrmy))
    u, cx, cy = np.empty((M + 1, B)), np.empty((M + 3, B)), np.empty((M + 3, B))
    st = np.zeros(B, dtype=np.int32)
    lib().hostsim_fit(_p(cenx), _p(ceny), _p(nrmx), _p(nrmy), _p(o), None, None, M, B, B, _p(u), _p(cx), _p(cy),
                      st.ctypes.data_as(_ip), int(split))
    return u.T.copy(), cx.T.copy(), cy.T.copy(), st


def fit_lsq(points, t, k):
    """points[B, M, 2], shared knots t[nt], degree k -> u[B, M+1], cx[B, nt-k-1], cy[B, nt-k-1], status[B]"""
    pts = np.asarray(points, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    B, M, _ = pts.shape
    nc = len(t) - k - 1
    px, py = sm(pts[:, :, 0]), sm(pts[:, :, 1])
    u, cx, cy = np.empty((M + 1, B)), np.empty((nc, B)), np.empty((nc, B))
    st = np.zeros(B, dtype=np.int32)
    rc = lib().hostsim_fit_lsq(_p(px), _p(py), M, B, B, _p(t), len(t), int(k), _p(u), _p(cx), _p(cy),
                               st.ctypes.data_as(_ip))
    if rc != 0:
        raise ValueError("unsupported sizes for the least-squares fit")
    return u.T.copy(), cx.T.copy(), cy.T.copy(), st


def evaluate(u, cx, cy, ts, split=1):
    u, cx, cy = sm(u), sm(cx), sm(cy)
    M, B = u.shape[0] - 1, u.shape[1]
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    N = len(ts)
    outs = [np.full((N, B), np.nan) for _ in range(6)]
    if split == 1:
        lib().hostsim_eval(_p(u), _p(cx), _p(cy), M, _p(ts), N, B, B, *[_p(o) for o in outs])
    else:
        lib().hostsim_eval_split(_p(u), _p(cx), _p(cy), M, _p(ts), N, B, B, int(split), *[_p(o) for o in outs])
    return dict(zip(("x", "y", "yaw", "radius", "chord_qss", "chord_norm"), [o.T.copy() for o in outs]))


def arc_sections(t, cx, cy, k, ts):
    t, cx, cy, ts = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, cx, cy, ts))
    sec = np.empty(len(ts))
    lib().hostsim_arc_sections(_p(t), len(t), _p(cx), _p(cy), int(k), _p(ts), len(ts), _p(sec))
    return sec


def eval_spline_batch(t, k, cx, cy, ts):
    """cx, cy: [B, n_coef] -> dict of [B, N] arrays"""
    t, ts = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, ts))
    cxs, cys = sm(cx), sm(cy)
    B, N = cxs.shape[1], len(ts)
    outs = [np.empty((N, B)) for _ in range(6)]
    lib().hostsim_eval_spline_batch(_p(t), len(t), int(k), _p(cxs), _p(cys), _p(ts), N, B, B, *[_p(o) for o in outs])
    return dict(zip(("x", "y", "yaw", "radius", "chord_qss", "chord_norm"), [o.T.copy() for o in outs]))


def eval_spline(t, cx, cy, k, ts):
    t, cx, cy, ts = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, cx, cy, ts))
    N = len(ts)
    outs = [np.empty(N) for _ in range(4)]
    lib().hostsim_eval_spline(_p(t), len(t), _p(cx), _p(cy), int(k), _p(ts), N, *[_p(o) for o in outs])
    return outs


def qss(impl, X, Y, R, sinb, veh, owner=False):
    """X, Y, R: [B, N].  Returns dict of [B, N] arrays + lap[B], summary[B, 8], status[B]."""
    x, y, r = sm(X), sm(Y), sm(R)
    N, B = x.shape
    sb = None if sinb is None else np.ascontiguousarray(sinb, dtype=np.float64)
    v, a, lat, tseg = (np.empty((N, B)) for _ in range(4))
    own = np.full((N, B), -1, dtype=np.int32) if owner else None
    lap, summ = np.empty(B), np.empty((8, B))
    st = np.zeros(B, dtype=np.int32)
    lib().hostsim_qss(int(impl), _p(x), _p(y), _p(r), _p(sb), N, B, B, C.byref(veh), _p(v), _p(a), _p(lat), _p(tseg),
                      None if own is None else own.ctypes.data_as(_ip), _p(lap), _p(summ), st.ctypes.data_as(_ip))
    return dict(v=v.T.copy(), a=a.T.copy(), lat=lat.T.copy(), time=tseg.T.copy(),
                owner=None if own is None else own.T.copy(), lap=lap, summary=summ.T.copy(), status=st)


def fast(cenx, ceny, nrmx, nrmy, sinb, ts, offsets, veh, rounds=2):
    """Fast mode (sto_fast.cuh) for offsets [B, M].  Returns dict(lap[B], status[B], v, a, time [B, N])."""
    off = sm(offsets)
    M, B = off.shape
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    N = len(ts)
    sb = None if sinb is None else np.ascontiguousarray(sinb, dtype=np.float64)
    c = [np.ascontiguousarray(x, dtype=np.float64) for x in (cenx, ceny, nrmx, nrmy)]
    v, a, tseg = (np.empty((N, B)) for _ in range(3))
    lap = np.empty(B)
    st = np.zeros(B, dtype=np.int32)
    lib().hostsim_fast(_p(c[0]), _p(c[1]), _p(c[2]), _p(c[3]), _p(sb), _p(ts), _p(off), M, N, B, B, C.byref(veh),
                       int(rounds), _p(v), _p(a), _p(tseg), _p(lap), st.ctypes.data_as(_ip))
    return dict(lap=lap, status=st, v=v.T.copy(), a=a.T.copy(), time=tseg.T.copy())
