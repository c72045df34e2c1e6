"""CPU: pins the C restatement (oracle/) against the golden vectors produced by the unmodified reference."""
import numpy as np
import pytest

import oracle_py as O
from helpers import CAND_CASES, SIM_CASES, golden, rel_err, veh_args


@pytest.mark.parametrize("name", SIM_CASES + ["sim_s30k5_i1"])
def test_qss_bit_exact_against_reference(name):
    """simulator.py:60-386 + fill_time: ref_pow mode reproduces SPEED/LON_ACC/LAT_ACC/ITERATION_FLAG/TIME bit for bit."""
    d = golden(name)
    veh = O.make_vehicle(*veh_args(d))
    r = O.qss(d["in_X"], d["in_Y"], d["in_CURVATURE"], np.sin(d["in_BANK"]), veh, ref_pow=1)
    assert r["status"] == 0
    assert np.array_equal(r["v"], d["out_SPEED"])
    assert np.array_equal(r["a"], d["out_LON_ACC"])
    assert np.array_equal(r["lat"], d["out_LAT_ACC"])
    assert np.array_equal(r["flag"], d["out_FLAG"])
    assert np.array_equal(r["time"], d["out_TIME"])
    assert abs(r["lap"] - float(d["lap"])) < 1e-11          # np.sum is pairwise, the oracle sums in order
    assert r["time"][0] == d["result_scalars"][0]            # total_time quirk = TIME[0]
    assert r["v"].max() == d["result_scalars"][2] and r["v"].min() == d["result_scalars"][3]


@pytest.mark.parametrize("name", SIM_CASES)
def test_qss_product_arithmetic_within_tolerance(name):
    """ref_pow=0 (x*x instead of libm pow(x, 2), what the GPU computes) stays within 1e-12 relative in speed."""
    d = golden(name)
    veh = O.make_vehicle(*veh_args(d))
    r = O.qss(d["in_X"], d["in_Y"], d["in_CURVATURE"], np.sin(d["in_BANK"]), veh, ref_pow=0)
    assert rel_err(r["v"], d["out_SPEED"]) < 1e-12
    assert np.array_equal(r["flag"], d["out_FLAG"])
    assert abs(r["lap"] - float(d["lap"])) < 1e-9


def test_synthetic_tables_inf_radius_and_tiny_n():
    d = golden("sim_synthetic_tables")
    veh = O.make_vehicle(*veh_args(d))
    for tag in ("n8", "n64", "n257"):
        r = O.qss(d[tag + "_in_X"], d[tag + "_in_Y"], d[tag + "_in_CURVATURE"], np.sin(d[tag + "_in_BANK"]), veh, 1)
        assert np.array_equal(r["v"], d[tag + "_out_SPEED"]) and np.array_equal(r["a"], d[tag + "_out_LON_ACC"])
        assert np.array_equal(r["time"], d[tag + "_out_TIME"]) and np.array_equal(r["flag"], d[tag + "_out_FLAG"])


def test_ppoly_and_ellipse_bit_exact():
    d = golden("sim_s10k3_i10_vehicle5")
    veh = O.make_vehicle(*veh_args(d))
    acc = [O.ppoly(d["veh_acc_x"], d["veh_acc_c"], g) for g in d["ppoly_grid"]]
    dcc = [O.ppoly(d["veh_dcc_x"], d["veh_dcc_c"], g) for g in d["ppoly_grid"]]
    assert np.array_equal(acc, d["ppoly_acc"]) and np.array_equal(dcc, d["ppoly_dcc"])
    lat = [O.maxlat(veh, l, 1) for l in d["ellipse_lon"]]
    assert np.array_equal(lat, d["ellipse_lat"])


@pytest.mark.parametrize("name", [c for c in SIM_CASES if "vehicle5" not in c])
def test_eval_deboor_order(name):
    """trajectory.py:250-260,278-281 through scipy _deBoor_D: X, Y bit-exact; radius within 4 ulp (NumPy's array
    pow is not libm's), yaw within 1e-15 (NumPy arctan2 vs libm)."""
    d = golden(name)
    X, Y, YAW, R = O.sample(d["spl_t"], d["spl_cx"], d["spl_cy"], int(d["spl_k"]), d["ts"], 1)
    assert np.array_equal(X, d["in_X"]) and np.array_equal(Y, d["in_Y"])
    assert rel_err(R, d["in_CURVATURE"]) < 1e-15
    assert np.max(np.abs(YAW - d["in_YAW"])) < 1e-15


@pytest.mark.parametrize("name", CAND_CASES)
def test_fit_against_fitpack(name):
    """trajectory.py:213-223 (splprep s=0,k=3,per=True).  The oracle's default solver restates FITPACK's fpclos Givens
    sweep: knots AND coefficients bit for bit, hence samples, speeds and the lap bit for bit end to end.  The two
    approximate solvers: coefficients within 1e-9 relative (BASELINE tolerance; measured ~1e-14), lap within 1e-6 s."""
    d = golden(name)
    veh = O.make_vehicle(*veh_args(d))
    for b in range(d["points"].shape[0]):
        O.set_fit_solver("fitpack")
        t, cx, cy = O.fit_periodic_cubic(d["points"][b])
        assert np.array_equal(t, d["ref_t"][b])
        assert np.array_equal(cx, d["ref_cx"][b]) and np.array_equal(cy, d["ref_cy"][b])
        X, Y, YAW, R = O.sample(t, cx, cy, 3, d["ts"], 1)
        assert np.array_equal(X, d["ref_X"][b]) and rel_err(R, d["ref_CURVATURE"][b]) < 1e-15
        r = O.qss(d["ref_X"][b], d["ref_Y"][b], d["ref_CURVATURE"][b], np.zeros(len(X)), veh, 1)
        assert np.array_equal(r["v"], d["ref_SPEED"][b]) and np.array_equal(r["time"], d["ref_TIME"][b])
        # end to end from the points: the oracle's own radius differs from NumPy's `** 3` by < 1 ulp, nothing else does
        r = O.qss(X, Y, R, np.zeros(len(X)), veh, 1)
        assert abs(r["lap"] - d["ref_lap"][b]) < 1e-10 and rel_err(r["v"], d["ref_SPEED"][b]) < 1e-12
        for solver in ("blocks", "thomas"):
            O.set_fit_solver(solver)
            try:
                t2, cx2, cy2 = O.fit_periodic_cubic(d["points"][b])
            finally:
                O.set_fit_solver("fitpack")
            assert np.array_equal(t2, t) and rel_err(cx2, cx) < 1e-9 and rel_err(cy2, cy) < 1e-9
            # speeds deviate up to ~3e-8 relative (conditioning of x''), laps stay within 1e-6 s here
            X2, Y2, YAW2, R2 = O.sample(t2, cx2, cy2, 3, d["ts"], 1)
            r2 = O.qss(X2, Y2, R2, np.zeros(len(X2)), veh, 1)
            assert abs(r2["lap"] - d["ref_lap"][b]) < 1e-6
            assert rel_err(r2["v"], d["ref_SPEED"][b]) < 1e-6


def test_lap_batch_threads_agree():
    d = golden("cand_m579_n579")
    veh = O.make_vehicle(*veh_args(d))
    nx, ny = -np.sin(d["centre_yaw"]), np.cos(d["centre_yaw"])
    l1, s1 = O.lap_batch(d["centre_x"], d["centre_y"], nx, ny, d["offsets"], d["ts"], np.zeros(len(d["ts"])), veh, 1, 1)
    l4, s4 = O.lap_batch(d["centre_x"], d["centre_y"], nx, ny, d["offsets"], d["ts"], np.zeros(len(d["ts"])), veh, 4, 1)
    assert np.array_equal(l1, l4) and not s1.any() and not s4.any()
    assert np.max(np.abs(l1 - d["ref_lap"])) < 1e-6


def test_vehicle_mirror_matches_reference_lookups():
    """Host Vehicle mirror: acc/dcc table look-ups and the friction-ellipse coupling, bit for bit against values the
    reference's Vehicle produced (models/vehicle.py:26-47), for the 5-row-table vehicle."""
    from spline_trajectory_optimization_b200.models.vehicle import Vehicle, VehicleParams
    d = golden("sim_s10k3_i10_vehicle5")
    v = Vehicle(VehicleParams(d["veh_acc_lookup"], d["veh_dcc_lookup"], *d["veh_scalars"]))
    assert np.array_equal(v.acc_intp.c, d["veh_acc_c"]) and np.array_equal(v.dcc_intp.x, d["veh_dcc_x"])
    assert np.array_equal([float(v.lookup_acc_from_speed(g)) for g in d["ppoly_grid"]], d["ppoly_acc"])
    assert np.array_equal([float(v.lookup_dcc_from_speed(g)) for g in d["ppoly_grid"]], d["ppoly_dcc"])
    assert np.array_equal([float(v.lookup_acc_circle(lon=l)[0]) for l in d["ellipse_lon"]], d["ellipse_lat"])
    pos, neg = v.lookup_acc_circle(lat=3.0)
    assert pos > 0 > neg
