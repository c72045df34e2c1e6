"""Shared test helpers: golden loading, vehicle construction for oracle / hostsim / product, GPU glue."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

SIM_CASES = ["sim_s30k5_i10", "sim_s30k5_i5", "sim_s30k5_i3", "sim_s10k3_i10", "sim_s10k3_i5", "sim_s10k3_i2",
             "sim_s0k3_i10", "sim_oval_bank12", "sim_oval_bank0", "sim_s10k3_i10_vehicle5"]
CAND_CASES = ["cand_m579_n579", "cand_m579_n1158", "cand_m2895_n2895"]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def veh_args(d):
    return (d["veh_scalars"], d["veh_acc_x"], d["veh_acc_c"], d["veh_dcc_x"], d["veh_dcc_c"])


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)))


def test_vehicle_params():
    from spline_trajectory_optimization_b200.models.vehicle import VehicleParams
    acc = np.array([[0.0, 10.0], [50.0, 7.0], [100.0, 0.5]])
    dcc = np.array([[0.0, -13.0], [50.0, -15.0], [100.0, -20.0]])
    return VehicleParams(acc, dcc, 10.0, -20.0, 15.0, -15.0, 100.0, 30.0)


test_vehicle_params.__test__ = False


def knots_from_u(u):
    """FITPACK's periodic knot vector from the chord-length parameter u[0..M] (u[M] = 1): three knots wrapped on each side."""
    u = np.asarray(u, dtype=np.float64)
    M = len(u) - 1
    return np.concatenate([u[M - 3:M] - 1.0, u, u[1:4] + 1.0])


def oracle_lap_from_coefficients(O, u, cx, cy, ts, sinb, veh, ref_pow=0):
    """Lap time of one line through the oracle's sampler + QSS + fill_time, starting from GIVEN spline coefficients
    (whatever solver produced them): the bit-exact check of everything downstream of the fit."""
    X, Y, YAW, R = O.sample(knots_from_u(u), cx, cy, 3, ts, ref_pow)
    sb = np.zeros(len(ts)) if sinb is None else np.asarray(sinb, dtype=np.float64)
    return O.qss(X, Y, R, sb, veh, ref_pow)["lap"]


def to_sm(a_cm, device="cuda"):
    """[B, n] host candidate-major -> [n, round_up(B, 32)] device sample-major (zero padded)."""
    import torch
    a = np.asarray(a_cm, dtype=np.float64)
    B, n = a.shape
    ld = (B + 31) & ~31
    out = torch.zeros((n, ld), dtype=torch.float64, device=device)
    out[:, :B] = torch.from_numpy(np.ascontiguousarray(a.T)).to(device)
    return out


def to_cm(t_sm, B):
    return t_sm[:, :B].T.contiguous().cpu().numpy()


def synthetic_closed_track(seed, n):
    """A closed line at ~2 m spacing from a piecewise-constant curvature (straights, sweepers, hairpins down to R = 8 m,
    optionally banked): many braking zones, re-initialisations and re-spawned rows on a few hundred samples."""
    rng = np.random.default_rng(seed)
    kappa = np.zeros(n)
    i = 0
    while i < n:
        L = int(rng.integers(5, 60))
        kind = rng.random()
        k = 0.0 if kind < 0.35 else (rng.normal(0, 0.01) if kind < 0.8 else rng.choice([-1, 1]) * rng.uniform(0.03, 0.12))
        kappa[i:i + L] = k
        i += L
    kappa += (2 * np.pi / (2.0 * n) - kappa.mean())
    psi = np.cumsum(kappa * 2.0)
    x, y = np.cumsum(2.0 * np.cos(psi)), np.cumsum(2.0 * np.sin(psi))
    with np.errstate(divide="ignore"):
        R = np.where(np.abs(kappa) > 1e-9, 1.0 / np.abs(kappa), np.inf)
    bank = rng.uniform(-0.05, 0.1, n) * (rng.random() < 0.5)
    return x, y, R, np.sin(bank)


def nan_cases():
    """NaN-producing inputs inside a live line (reference simulator.py:239: the spawn test `g > max_curve_speed or
    g < min_state_speed` is false for NaN, so an infeasible step on NaN operands only stops the front)."""
    x, y, r, sb = synthetic_closed_track(3, 300)
    cases = {}
    xr, yr, rr = x.copy(), y.copy(), r.copy()
    rr[120] = np.nan
    cases["nan_radius"] = (xr, yr, rr, sb)
    xp, yp, rp = x.copy(), y.copy(), r.copy()
    xp[57] = np.nan
    cases["nan_position"] = (xp, yp, rp, sb)
    cases["all_nan"] = (np.full_like(x, np.nan), np.full_like(y, np.nan), np.full_like(r, np.nan), sb)
    xm, ym, rm = x.copy(), y.copy(), r.copy()
    rm[10:14] = np.nan
    ym[200] = np.nan
    cases["nan_mixed"] = (xm, ym, rm, sb)
    return cases
