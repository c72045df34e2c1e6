"""GPU (-m gpu): the CUDA path, called through the C ABI (libsto_b200.so), against the oracle and the golden
vectors of the unmodified reference.

Tolerances (BASELINE.json north_star): spline coefficients and speed profiles within 1e-9 relative, lap time
within 1e-6 s on identical inputs.  What is actually asserted is tighter:
  * against the oracle in ref_pow=0 mode (the kernels' arithmetic): BIT-EXACT coefficients, samples, speeds,
    accelerations, segment times and laps;
  * against the reference's golden vectors on identical inputs: speeds <= 1e-12 relative (libm pow(x,2) vs x*x),
    laps <= 1e-9 s; end to end through our own fit: lap <= 1e-6 s (speeds deviate <= ~3e-8).
"""
import contextlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle_py as O  # noqa: E402
from helpers import (CAND_CASES, SIM_CASES, golden, oracle_lap_from_coefficients, rel_err,  # noqa: E402
                     test_vehicle_params, to_cm, to_sm,
                     veh_args)


@pytest.fixture(scope="module")
def sto():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from spline_trajectory_optimization_b200 import _lib, evaluator
    _lib.load()   # fails loudly if libsto_b200.so is missing: there is no fallback
    return evaluator


@contextlib.contextmanager
def fit_solver(name):
    """Selects the solver of the fit's interpolation system in the library AND in the oracle, which restates all three:
    'fitpack' = the default, FITPACK's own Givens sweep (bit-identical to the reference's SciPy fit), 'partitioned' = 32
    blocks + cyclic reduction (lane groups on the device), 'thomas' = one-lane Thomas + Sherman-Morrison recurrences.
    With matching solvers everything is bit-exact against O.lap_batch."""
    from spline_trajectory_optimization_b200 import _lib
    lib = _lib.load()
    code = {"thomas": _lib.FIT_THOMAS, "partitioned": _lib.FIT_BLOCKS, "fitpack": _lib.FIT_FITPACK}[name]
    _lib.check(lib.sto_set_fit_solver(code))
    O.set_fit_solver({"thomas": "thomas", "partitioned": "blocks", "fitpack": "fitpack"}[name])   # the oracle restates all three
    try:
        yield
    finally:
        _lib.check(lib.sto_set_fit_solver(_lib.FIT_FITPACK))
        O.set_fit_solver("fitpack")


def _evaluator(sto, d, ts=None, bank=None, impl="memo"):
    from spline_trajectory_optimization_b200 import _lib
    nrm = np.stack([-np.sin(d["centre_yaw"]), np.cos(d["centre_yaw"])], axis=1)
    ctr = np.stack([d["centre_x"], d["centre_y"]], axis=1)
    veh = _lib.make_vehicle(*veh_args(d))
    return sto.BatchedLineEvaluator(ctr, nrm, d["ts"] if ts is None else ts, veh, bank=bank, impl=impl)


@pytest.mark.parametrize("name", CAND_CASES)
@pytest.mark.parametrize("solver", ["fitpack", "thomas", "partitioned"])
def test_fit_and_sample(sto, name, solver):
    with fit_solver(solver):
        _fit_and_sample(sto, name, solver)


def _fit_and_sample(sto, name, solver):
    d = golden(name)
    ev = _evaluator(sto, d)
    B, M = d["offsets"].shape
    u, cx, cy, st = ev.fit(to_sm(d["offsets"]), B=B)
    S = ev.sample(u, cx, cy, B=B, want=("x", "y", "yaw", "radius", "chord_qss", "chord_norm"))
    torch.cuda.synchronize()
    assert not st[:B].any()
    u, cx, cy = to_cm(u, B), to_cm(cx, B), to_cm(cy, B)
    X, Y, YAW, R = (to_cm(S[k], B) for k in ("x", "y", "yaw", "radius"))
    for b in range(B):
        t, ocx, ocy = O.fit_periodic_cubic(d["points"][b])
        assert np.array_equal(u[b], t[3:-3]) and np.array_equal(cx[b], ocx) and np.array_equal(cy[b], ocy)
        assert np.array_equal(t, d["ref_t"][b])                       # knots: exact vs FITPACK
        assert rel_err(cx[b], d["ref_cx"][b]) < 1e-9 and rel_err(cy[b], d["ref_cy"][b]) < 1e-9
        if solver == "fitpack":                                       # the reference's own coefficients and samples
            assert np.array_equal(cx[b], d["ref_cx"][b]) and np.array_equal(cy[b], d["ref_cy"][b])
            assert np.array_equal(X[b], d["ref_X"][b]) and np.array_equal(Y[b], d["ref_Y"][b])
            assert rel_err(R[b], d["ref_CURVATURE"][b]) < 1e-15
        oX, oY, oYAW, oR = O.sample(t, ocx, ocy, 3, d["ts"], 0)
        assert np.array_equal(X[b], oX) and np.array_equal(Y[b], oY) and np.array_equal(R[b], oR)
        assert np.max(np.abs(YAW[b] - oYAW)) < 1e-14                  # device atan2 vs libm
        assert np.max(np.abs(X[b] - d["ref_X"][b])) < 1e-9            # end to end vs the reference's samples
    # fit from explicit points == fit from centre + offset * normal
    px, py = to_sm(d["points"][:, :, 0]), to_sm(d["points"][:, :, 1])
    u2, cx2, cy2, _ = ev.fit(points_sm=(px, py), B=B)
    assert np.array_equal(to_cm(cx2, B), cx) and np.array_equal(to_cm(u2, B), u)


@pytest.mark.parametrize("name", CAND_CASES)
def test_sample_on_reference_coefficients(sto, name):
    """Per stage on identical inputs: the reference's (FITPACK) coefficients through our sampler."""
    d = golden(name)
    ev = _evaluator(sto, d)
    B = d["offsets"].shape[0]
    u = to_sm(d["ref_t"][:, 3:-3])
    S = ev.sample(u, to_sm(d["ref_cx"]), to_sm(d["ref_cy"]), B=B)
    assert np.array_equal(to_cm(S["x"], B), d["ref_X"]) and np.array_equal(to_cm(S["y"], B), d["ref_Y"])
    assert rel_err(to_cm(S["radius"], B), d["ref_CURVATURE"]) < 1e-15
    assert np.max(np.abs(to_cm(S["yaw"], B) - d["ref_YAW"])) < 1e-14


@pytest.mark.parametrize("name", SIM_CASES + ["sim_s30k5_i1"])
@pytest.mark.parametrize("impl", ["plain", "memo"])
def test_qss_against_oracle_and_reference(sto, name, impl):
    from spline_trajectory_optimization_b200 import _lib
    d = golden(name)
    veh, ov = _lib.make_vehicle(*veh_args(d)), O.make_vehicle(*veh_args(d))
    sb = np.sin(d["in_BANK"])
    res = sto.run_qss(to_sm(d["in_X"][None]), to_sm(d["in_Y"][None]), to_sm(d["in_CURVATURE"][None]), veh, B=1,
                      sin_bank=sb, impl=sto.IMPL[impl], owner=(impl == "plain"))
    torch.cuda.synchronize()
    assert int(res["status"][0]) == 0
    o = O.qss(d["in_X"], d["in_Y"], d["in_CURVATURE"], sb, ov, 0)
    v, a, lat, tm = (res[k][:, 0].cpu().numpy() for k in ("speed", "lon_acc", "lat_acc", "time"))
    assert np.array_equal(v, o["v"]) and np.array_equal(a, o["a"])          # bit-exact vs the oracle
    assert np.array_equal(lat, o["lat"]) and np.array_equal(tm, o["time"])
    assert float(res["lap"][0]) == o["lap"]
    assert float(res["summary"][6, 0]) == o["steps"]
    if impl == "plain":
        assert np.array_equal(res["owner"][:, 0].cpu().numpy(), d["out_FLAG"].astype(np.int32))
    # vs the unmodified reference
    assert rel_err(v, d["out_SPEED"]) < 1e-12
    assert abs(float(res["lap"][0]) - float(d["lap"])) < 1e-9
    assert np.max(np.abs(tm - d["out_TIME"])) < 1e-13
    assert abs(float(res["summary"][0, 0]) - d["result_scalars"][0]) < 1e-13   # total_time quirk (= TIME[0])


@pytest.mark.gpu
def test_qss_memo_word_boundaries(sto):
    """Track sizes around multiples of 64 (plane word size), 8 different lines per size so that a warp holds lane groups
    with different lists: the memoised kernel (lane groups, vectorised list walk, dense iteration-0 sweep) against the
    oracle, bit for bit, including the front-step count."""
    from spline_trajectory_optimization_b200 import _lib
    d = golden("sim_s10k3_i2")
    veh, ov = _lib.make_vehicle(*veh_args(d)), O.make_vehicle(*veh_args(d))
    n_full = len(d["in_X"])
    for n in (128, 129, 191, 192, 193, 257, 1025):
        B = 8
        X, Y, R = np.empty((B, n)), np.empty((B, n)), np.empty((B, n))
        for c in range(B):
            idx = (np.linspace(0, n_full - 1, n, endpoint=False).astype(int) + 37 * c) % n_full
            X[c], Y[c], R[c] = d["in_X"][idx], d["in_Y"][idx], d["in_CURVATURE"][idx] * (1.0 + 0.03 * c)
        res = sto.run_qss(to_sm(X), to_sm(Y), to_sm(R), veh, B=B, sin_bank=np.zeros(n), impl=sto.IMPL["memo"])
        torch.cuda.synchronize()
        assert not res["status"][:B].cpu().numpy().any()
        for c in range(B):
            o = O.qss(X[c], Y[c], R[c], np.zeros(n), ov, 0)
            assert np.array_equal(res["speed"][:, c].cpu().numpy(), o["v"]), (n, c)
            assert np.array_equal(res["time"][:, c].cpu().numpy(), o["time"]), (n, c)
            assert float(res["lap"][c]) == o["lap"] and float(res["summary"][6, c]) == o["steps"], (n, c)


def test_qss_memo_random_tracks(sto):
    """Eight different random synthetic tracks per batch (hairpins, banked stretches), three sizes: a warp then holds lane
    groups with very different lists and round counts.  Memoised kernel against the oracle, bit for bit."""
    from helpers import synthetic_closed_track
    from spline_trajectory_optimization_b200 import _lib
    d = golden("sim_s10k3_i2")
    veh, ov = _lib.make_vehicle(*veh_args(d)), O.make_vehicle(*veh_args(d))
    for n in (200, 333, 640):
        B = 8
        T = [synthetic_closed_track(1000 * n + c, n) for c in range(B)]
        X, Y, R = (np.stack([t[k] for t in T]) for k in range(3))
        sb = T[0][3]                                            # the bank profile is shared by a batch
        res = sto.run_qss(to_sm(X), to_sm(Y), to_sm(R), veh, B=B, sin_bank=sb, impl=sto.IMPL["memo"])
        torch.cuda.synchronize()
        assert not res["status"][:B].cpu().numpy().any()
        for c in range(B):
            o = O.qss(X[c], Y[c], R[c], sb, ov, 0)
            assert np.array_equal(res["speed"][:, c].cpu().numpy(), o["v"]), (n, c)
            assert np.array_equal(res["time"][:, c].cpu().numpy(), o["time"]), (n, c)
            assert float(res["lap"][c]) == o["lap"] and float(res["summary"][6, c]) == o["steps"], (n, c)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [1, 2, 3, 4])
def test_qss_kernel_variants_equal_oracle(sto, kernel):
    """The selectable small-batch kernels (tuning key qss_kernel): 1 the four-walker lane-group kernel, 2 the one-loop
    kernel with the sequential forward sweep, 3 the one-loop kernel with the re-spawned lists walked out of order
    (sto_qss_memo3.cuh), 4 = default, the one-loop kernel with the forward original-row sub-pass run-parallel.  Random tracks
    (8 different lines per warp), word-boundary sizes, 16 / 32-lane groups and 64 lines of the golden Monza line with
    perturbed radii (4 x 8 lanes): speeds, segment times, laps, front-step and iteration counts equal the oracle's."""
    from helpers import synthetic_closed_track
    from spline_trajectory_optimization_b200 import _lib
    lib = _lib.load()
    d = golden("sim_s10k3_i2")
    veh, ov = _lib.make_vehicle(*veh_args(d)), O.make_vehicle(*veh_args(d))
    n_full = len(d["in_X"])
    batches = []
    for n in (200, 640):
        T = [synthetic_closed_track(77000 + 10 * n + c, n) for c in range(8)]
        batches.append(tuple(np.stack([t[k] for t in T]) for k in range(3)) + (T[0][3], 0))
    for n in (129, 193, 1025):
        X, Y, R = np.empty((8, n)), np.empty((8, n)), np.empty((8, n))
        for c in range(8):
            idx = (np.linspace(0, n_full - 1, n, endpoint=False).astype(int) + 53 * c) % n_full
            X[c], Y[c], R[c] = d["in_X"][idx], d["in_Y"][idx], d["in_CURVATURE"][idx] * (1.0 + 0.02 * c)
        batches.append((X, Y, R, np.zeros(n), 0))
    B = 64
    X = np.tile(d["in_X"], (B, 1)); Y = np.tile(d["in_Y"], (B, 1))
    R = d["in_CURVATURE"][None, :] * (1.0 + 0.004 * np.arange(B))[:, None]
    batches.append((X, Y, R, np.sin(d["in_BANK"]), 8))       # forced 4 lines x 8 lanes per warp
    batches.append((X[:6], Y[:6], R[:6], np.sin(d["in_BANK"]), 16))
    try:
        _lib.check(lib.sto_set_tuning(b"qss_kernel", kernel))
        for X, Y, R, sb, group in batches:
            _lib.check(lib.sto_set_tuning(b"qss_group", group))
            nb = X.shape[0]
            res = sto.run_qss(to_sm(X), to_sm(Y), to_sm(R), veh, B=nb, sin_bank=sb, impl=sto.IMPL["memo"])
            torch.cuda.synchronize()
            assert not res["status"][:nb].cpu().numpy().any()
            for c in range(0, nb, 1 if nb <= 8 else 9):
                o = O.qss(X[c], Y[c], R[c], sb, ov, 0)
                assert np.array_equal(res["speed"][:, c].cpu().numpy(), o["v"]), (kernel, X.shape, c)
                assert np.array_equal(res["lon_acc"][:, c].cpu().numpy(), o["a"]), (kernel, X.shape, c)
                assert np.array_equal(res["time"][:, c].cpu().numpy(), o["time"]), (kernel, X.shape, c)
                assert float(res["lap"][c]) == o["lap"], (kernel, X.shape, c)
                assert float(res["summary"][6, c]) == o["steps"] and float(res["summary"][7, c]) == o["iters"] + 1
    finally:
        _lib.check(lib.sto_set_tuning(b"qss_kernel", 0))
        _lib.check(lib.sto_set_tuning(b"qss_group", 0))


@pytest.mark.parametrize("impl", ["plain", "memo"])
def test_qss_synthetic_tables(sto, impl):
    """Infinite turn radius, banked samples, N = 8 / 64 / 257 (tiny N runs the plain kernel under 'memo')."""
    from spline_trajectory_optimization_b200 import _lib
    d = golden("sim_synthetic_tables")
    veh = _lib.make_vehicle(*veh_args(d))
    for tag in ("n8", "n64", "n257"):
        res = sto.run_qss(to_sm(d[tag + "_in_X"][None]), to_sm(d[tag + "_in_Y"][None]),
                          to_sm(d[tag + "_in_CURVATURE"][None]), veh, B=1, sin_bank=np.sin(d[tag + "_in_BANK"]),
                          impl=sto.IMPL[impl])
        assert rel_err(res["speed"][:, 0].cpu().numpy(), d[tag + "_out_SPEED"]) < 1e-12
        assert np.max(np.abs(res["time"][:, 0].cpu().numpy() - d[tag + "_out_TIME"])) < 1e-12


@pytest.mark.parametrize("name", CAND_CASES)
@pytest.mark.parametrize("impl", ["plain", "memo"])
@pytest.mark.parametrize("solver", ["fitpack", "thomas", "partitioned"])
def test_fused_lap_time(sto, name, impl, solver):
    d = golden(name)
    ev = _evaluator(sto, d, impl=impl)
    B = d["offsets"].shape[0]
    ov = O.make_vehicle(*veh_args(d))
    nx, ny = -np.sin(d["centre_yaw"]), np.cos(d["centre_yaw"])
    olap, ost = O.lap_batch(d["centre_x"], d["centre_y"], nx, ny, d["offsets"], d["ts"], np.zeros(len(d["ts"])),
                            ov, n_threads=4, ref_pow=0)
    with fit_solver(solver):
        lap, st = ev.lap_times(to_sm(d["offsets"]), B=B)
        lap, st = lap.cpu().numpy(), st.cpu().numpy()
        assert not st.any()
        olap_s, _ = O.lap_batch(d["centre_x"], d["centre_y"], nx, ny, d["offsets"], d["ts"], np.zeros(len(d["ts"])),
                                ov, n_threads=4, ref_pow=0)
        assert np.array_equal(lap, olap_s)                               # bit-exact vs the oracle, fit included
        assert np.max(np.abs(lap - olap)) < 1e-7                         # the two solvers: last bits of c only
        u, cx, cy, _ = ev.fit(to_sm(d["offsets"]), B=B)                 # downstream of the fit from GIVEN coefficients
        u, cx, cy = to_cm(u, B), to_cm(cx, B), to_cm(cy, B)
        for b in range(B):
            assert lap[b] == oracle_lap_from_coefficients(O, u[b], cx[b], cy[b], d["ts"], None, ov)
        assert np.max(np.abs(lap - d["ref_lap"])) < 1e-6                  # BASELINE: lap within 1e-6 s of the reference
        if solver == "fitpack":                                          # reference coefficients: only x*x vs pow(x, 2) is left
            assert np.max(np.abs(lap - d["ref_lap"])) < 1e-10
        # host-buffer entry point (H2D + transpose + D2H inside) returns the same bits
        hlap, hst = ev.lap_times_host(d["offsets"])
        assert np.array_equal(hlap, lap) and not hst.any()


def test_simulator_drop_in(sto):
    """The reference-facing API: Simulator(vehicle).run_simulation(traj) on the reference's own sampled table."""
    from spline_traj_optm.models.trajectory import Trajectory
    from spline_traj_optm.models.vehicle import Vehicle
    from spline_traj_optm.simulator.simulator import Simulator
    d = golden("sim_s30k5_i10")
    traj = Trajectory(len(d["in_X"]))
    for c, k in ((Trajectory.X, "in_X"), (Trajectory.Y, "in_Y"), (Trajectory.YAW, "in_YAW"),
                 (Trajectory.CURVATURE, "in_CURVATURE"), (Trajectory.DIST_TO_SF_BWD, "in_DIST_BWD"),
                 (Trajectory.DIST_TO_SF_FWD, "in_DIST_FWD"), (Trajectory.BANK, "in_BANK")):
        traj[:, c] = d[k]
    before = traj.points.copy()
    res = Simulator(Vehicle(test_vehicle_params()), track_iteration_flag=True).run_simulation(traj, False)
    assert np.array_equal(traj.points, before)                        # input untouched (simulator.py:64)
    out = res.trajectory
    assert rel_err(out[:, Trajectory.SPEED], d["out_SPEED"]) < 1e-12
    assert rel_err(out[:, Trajectory.LAT_ACC], d["out_LAT_ACC"]) < 1e-11
    assert np.max(np.abs(out[:, Trajectory.LON_ACC] - d["out_LON_ACC"])) < 1e-11
    assert np.array_equal(out[:, Trajectory.ITERATION_FLAG], d["out_FLAG"])
    assert np.max(np.abs(out[:, Trajectory.TIME] - d["out_TIME"])) < 1e-13
    ref = d["result_scalars"]
    got = [res.total_time, res.average_speed, res.max_speed, res.min_speed, res.max_lat_acc, res.max_lon_acc,
           res.max_lon_dcc]
    assert np.allclose(got, ref, rtol=1e-11, atol=0)
    assert abs(res.lap_time - float(d["lap"])) < 1e-9
    assert "Lap Time" in str(res)
    # the default (memoised kernel): same profile bit for bit, ITERATION_FLAG (debug column) not tracked
    fast = Simulator(Vehicle(test_vehicle_params())).run_simulation(traj, False)
    assert np.allclose([fast.total_time, fast.average_speed, fast.max_speed, fast.min_speed, fast.max_lat_acc,
                        fast.max_lon_acc, fast.max_lon_dcc], ref, rtol=1e-11, atol=0)
    for c in (Trajectory.SPEED, Trajectory.LON_ACC, Trajectory.LAT_ACC, Trajectory.TIME):
        assert np.array_equal(fast.trajectory[:, c], out[:, c])
    assert (fast.trajectory[:, Trajectory.ITERATION_FLAG] == -1).all() and fast.lap_time == res.lap_time


def test_bspline_trajectory_sample_along_on_device(sto):
    """BSplineTrajectory(points, s=30, k=5).sample_along(10.0): SciPy fit on the host, GPU evaluation."""
    from spline_traj_optm.models.trajectory import BSplineTrajectory, Trajectory
    from spline_trajectory_optimization_b200 import tracks
    d = golden("sim_s30k5_i10")
    spl = BSplineTrajectory(tracks.monza_raw()[0], 30.0, 5)
    assert np.array_equal(spl._spl_x.c, d["spl_cx"]) and spl.get_length() == float(d["spl_length"])
    traj = spl.sample_along(10.0)
    assert np.array_equal(traj[:, Trajectory.X], d["in_X"]) and np.array_equal(traj[:, Trajectory.Y], d["in_Y"])
    assert rel_err(traj[:, Trajectory.CURVATURE], d["in_CURVATURE"]) < 1e-15
    assert np.max(np.abs(traj[:, Trajectory.YAW] - d["in_YAW"])) < 1e-14
    assert np.allclose(traj[:, Trajectory.DIST_TO_SF_BWD], d["in_DIST_BWD"], rtol=1e-12)
    # the GPU's fixed-order quadrature for the same columns (models/trajectory.py:283-289)
    fast = spl.sample_along(10.0, arc_length="gauss")
    assert np.allclose(fast[1:, Trajectory.DIST_TO_SF_BWD], d["in_DIST_BWD"][1:], rtol=1e-9, atol=0)
    assert np.allclose(fast[:, Trajectory.DIST_TO_SF_FWD], d["in_DIST_FWD"], rtol=0, atol=1e-6)   # metres of 5.8 km
    assert np.array_equal(fast[:, Trajectory.X], d["in_X"])


def test_edge_cases(sto):
    """B = 1, ragged B, minimal M = 3, degenerate (zero-length) line, argmin."""
    from spline_trajectory_optimization_b200 import _lib
    d = golden("cand_m579_n579")
    ev = _evaluator(sto, d)
    lap_all, _ = ev.lap_times(to_sm(d["offsets"]), B=8)
    lap1, st1 = ev.lap_times(to_sm(d["offsets"][3:4]), B=1)
    assert float(lap1[0]) == float(lap_all[3]) and int(st1[0]) == 0
    lap5, _ = ev.lap_times(to_sm(d["offsets"][:5]), B=5)             # ragged: 5 of a 32-wide warp
    assert torch.equal(lap5, lap_all[:5])
    best, idx = ev.argmin(lap_all.contiguous())
    assert int(idx[0]) == int(np.argmin(lap_all.cpu().numpy())) and float(best[0]) == float(lap_all.min())
    # M = 3: the smallest closed line FITPACK accepts (trajectory.py:214)
    tri = np.array([[[0.0, 0.0], [40.0, 0.0], [20.0, 30.0]]])
    z = np.zeros((1, 5))
    for solver in ("fitpack", "thomas", "partitioned"):
        with fit_solver(solver):
            t, ocx, ocy = O.fit_periodic_cubic(tri[0])
            u, cx, cy, st = ev.fit(points_sm=(to_sm(tri[:, :, 0]), to_sm(tri[:, :, 1])), B=1)
            assert int(st[0]) == 0 and np.array_equal(to_cm(u, 1)[0], t[3:-3])
            assert np.array_equal(to_cm(cx, 1)[0], ocx) and np.array_equal(to_cm(cy, 1)[0], ocy)
            # all points identical -> FITPACK would refuse (ier=10); we flag the candidate instead of aborting the batch
            u, cx, cy, st = ev.fit(points_sm=(to_sm(z), to_sm(z)), B=1)
            assert int(st[0]) & _lib.CAND_DEGENERATE_FIT and bool(torch.isnan(cx[:, 0]).all())


def test_full_size_properties(sto):
    """BASELINE config 2 size (Monza, M = N = 2895, 4,096 candidates): size-independent properties."""
    from spline_trajectory_optimization_b200 import candidates, tracks
    from spline_trajectory_optimization_b200.models.race_track import RaceTrack
    from spline_trajectory_optimization_b200.models.vehicle import Vehicle
    c, l, r = tracks.monza_raw()
    rt = RaceTrack("monza", l, r, c, s=10.0, interval=2.0)
    M = len(rt.center_d)
    assert M == 2895
    B = 4096
    off = candidates.smooth_offsets(M, B, rt.dist_to_left, rt.dist_to_right, seed=1234)
    ev = sto.BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), Vehicle(test_vehicle_params()))
    d_off = ev.to_sample_major(torch.from_numpy(off).cuda())
    lap, st = ev.lap_times(d_off, B=B)
    lap, st = lap.cpu().numpy(), st.cpu().numpy()
    assert not st.any() and np.isfinite(lap).all()
    # candidate 0 is the centre line: equals the single-line reference result (golden config-2 candidate 0)
    g = golden("cand_m2895_n2895")
    assert abs(lap[0] - g["ref_lap"][0]) < 1e-6
    assert 100.0 < lap.min() and lap.max() < 120.0
    # oracle spot checks, bit-exact (fit included; the same lines fitted in a batch of 4 use 32 lanes each instead of 8
    # and still give the same bits), with either solver of the collocation system
    ov = O.make_vehicle(*veh_args(g))
    nrm = rt.left_normals()
    pick = [0, 1, 777, 4095]
    olap, _ = O.lap_batch(rt.center_d[:, 0], rt.center_d[:, 1], nrm[:, 0], nrm[:, 1], off[pick], rt.center_d.ts(),
                          np.zeros(M), ov, n_threads=4, ref_pow=0)
    assert np.array_equal(lap[pick], olap)
    lap4, _ = ev.lap_times(to_sm(off[pick]), B=4)
    assert np.array_equal(lap4.cpu().numpy(), olap)
    with fit_solver("thomas"):
        lap_t, _ = ev.lap_times(d_off, B=B)
        olap_t, _ = O.lap_batch(rt.center_d[:, 0], rt.center_d[:, 1], nrm[:, 0], nrm[:, 1], off[pick], rt.center_d.ts(),
                                np.zeros(M), ov, n_threads=4, ref_pow=0)
    assert np.array_equal(lap_t.cpu().numpy()[pick], olap_t)
    assert np.max(np.abs(olap_t - olap)) < 1e-7
    # permutation invariance: candidates are independent (what the multi-GPU sharding relies on)
    perm = np.random.default_rng(0).permutation(B)
    lap_p, _ = ev.lap_times(ev.to_sample_major(torch.from_numpy(off[perm]).cuda()), B=B)
    assert np.array_equal(lap_p.cpu().numpy(), lap[perm])
    # the two schedule kernels agree bit for bit on a 256-candidate slice
    evp = sto.BatchedLineEvaluator(rt.center_d[:, :2], nrm, rt.center_d.ts(), Vehicle(test_vehicle_params()), impl="plain")
    lap_plain, _ = evp.lap_times(evp.to_sample_major(torch.from_numpy(off[:256]).cuda()), B=256)
    assert np.array_equal(lap_plain.cpu().numpy(), lap[:256])


def test_full_size_parity_512_candidates(sto):
    """BASELINE configs[1] at its stated size (M = N = 2895, a 4,096-candidate launch): every 8th candidate - 512 lines -
    against the oracle.  ref_pow = 0 (the kernels' x*x arithmetic): laps BIT-EXACT.  ref_pow = 1 (the reference's libm
    pow(x, 2), the arithmetic the goldens pin bit for bit): |lap difference| <= 1e-6 s, BASELINE's lap tolerance, asserted
    (measured ~2.5e-10 s)."""
    import os
    from spline_trajectory_optimization_b200 import candidates, tracks
    from spline_trajectory_optimization_b200.models.race_track import RaceTrack
    from spline_trajectory_optimization_b200.models.vehicle import Vehicle
    c, l, r = tracks.monza_raw()
    rt = RaceTrack("monza", l, r, c, s=10.0, interval=2.0)
    M, B = len(rt.center_d), 4096
    off = candidates.smooth_offsets(M, B, rt.dist_to_left, rt.dist_to_right, seed=4321)
    ev = sto.BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), Vehicle(test_vehicle_params()))
    lap, st = ev.lap_times(ev.to_sample_major(torch.from_numpy(off).cuda()), B=B)
    lap, st = lap.cpu().numpy(), st.cpu().numpy()
    assert not st.any()
    pick = np.arange(0, B, 8)
    assert len(pick) == 512
    g = golden("cand_m2895_n2895")
    ov = O.make_vehicle(*veh_args(g))
    nrm = rt.left_normals()
    threads = min(32, os.cpu_count() or 1)
    args = (rt.center_d[:, 0], rt.center_d[:, 1], nrm[:, 0], nrm[:, 1], off[pick], rt.center_d.ts(), np.zeros(M), ov)
    olap0, ost0 = O.lap_batch(*args, n_threads=threads, ref_pow=0)
    assert not ost0.any() and np.array_equal(lap[pick], olap0)
    olap1, ost1 = O.lap_batch(*args, n_threads=threads, ref_pow=1)
    assert not ost1.any()
    assert np.max(np.abs(lap[pick] - olap1)) <= 1e-6, np.max(np.abs(lap[pick] - olap1))


def test_fused_banked_oval(sto):
    """BASELINE config 4 (bank-aware QSS through the fused path): synthetic banked oval, 4-column centre line, bank
    carried to the samples by index; bit-exact against the oracle."""
    from spline_trajectory_optimization_b200 import candidates
    from spline_trajectory_optimization_b200.models.race_track import RaceTrack
    from spline_trajectory_optimization_b200.models.trajectory import Trajectory
    from spline_trajectory_optimization_b200.models.vehicle import Vehicle
    centre, left, right = candidates.banked_oval(straight=300.0, radius=120.0, spacing=10.0)
    rt = RaceTrack("oval", left, right, centre, s=1.0, interval=4.0)
    M = len(rt.center_d)
    bank = rt.center_d[:, Trajectory.BANK]
    assert bank.max() > 0.1 and bank.min() >= 0.0
    B = 37
    off = candidates.smooth_offsets(M, B, rt.dist_to_left, rt.dist_to_right, seed=5)
    veh = Vehicle(test_vehicle_params())
    nrm = rt.left_normals()
    laps = {}
    for impl in ("plain", "memo"):
        ev = sto.BatchedLineEvaluator(rt.center_d[:, :2], nrm, rt.center_d.ts(), veh, bank=bank, impl=impl)
        lap, st = ev.lap_times(to_sm(off), B=B)
        assert not st.cpu().numpy().any()
        laps[impl] = lap.cpu().numpy()
        with fit_solver("thomas"):
            lap_t, st_t = ev.lap_times(to_sm(off), B=B)
        assert not st_t.cpu().numpy().any() and np.max(np.abs(lap_t.cpu().numpy() - laps[impl])) < 1e-7
    g = golden("cand_m579_n579")
    ov = O.make_vehicle(*veh_args(g))
    olap, ost = O.lap_batch(rt.center_d[:, 0], rt.center_d[:, 1], nrm[:, 0], nrm[:, 1], off, rt.center_d.ts(),
                            np.sin(bank), ov, n_threads=4, ref_pow=0)
    assert np.array_equal(laps["plain"], olap) and np.array_equal(laps["memo"], olap)
    # banking matters: the same lines on the flat track are faster (reference sign quirk)
    ev0 = sto.BatchedLineEvaluator(rt.center_d[:, :2], nrm, rt.center_d.ts(), veh, bank=None)
    lap0, _ = ev0.lap_times(to_sm(off), B=B)
    assert (lap0.cpu().numpy() < laps["memo"]).all()


def test_banked_oval_of_the_baseline_config(sto):
    """BASELINE configs[3] / SURVEY.md 8(d): two 800 m straights + two R = 250 m semicircles, width 15 m, bank 0 -> 9 deg
    blended over 100 m, M = N at 2 m (the track bench.py --config 3 runs): 96 candidates, bit-exact against the oracle."""
    import bench
    from spline_trajectory_optimization_b200 import candidates
    from spline_trajectory_optimization_b200.models.vehicle import Vehicle
    rt = bench.build_track(2.0, "oval")
    M = len(rt.center_d)
    assert 1500 < M < 1700
    bank = bench.track_bank(rt)
    assert bank is not None and abs(np.rad2deg(bank.max()) - 9.0) < 0.2 and bank.min() >= 0.0
    B = 96
    off = candidates.smooth_offsets(M, B, rt.dist_to_left, rt.dist_to_right, seed=9)
    assert np.abs(off).max() > 3.0                                # the 15 m track leaves +-6.5 m
    nrm = rt.left_normals()
    ev = sto.BatchedLineEvaluator(rt.center_d[:, :2], nrm, rt.center_d.ts(), Vehicle(test_vehicle_params()), bank=bank)
    lap, st = ev.lap_times(to_sm(off), B=B)
    assert not st.cpu().numpy().any()
    g = golden("cand_m579_n579")
    ov = O.make_vehicle(*veh_args(g))
    olap, ost = O.lap_batch(rt.center_d[:, 0], rt.center_d[:, 1], nrm[:, 0], nrm[:, 1], off, rt.center_d.ts(),
                            np.sin(bank), ov, n_threads=8, ref_pow=0)
    assert not ost.any() and np.array_equal(lap.cpu().numpy(), olap)


def test_long_track_quarter_metre(sto):
    """BASELINE config 5 geometry (Monza at 0.25 m: M = N = 23,160; ~25 M front steps per line in the reference): four
    candidates through the fused path, bit-exact against the oracle.  Exercises multi-word rings (362 words), the
    reduced candidates-per-warp launch and a 23 k-row Thomas solve."""
    from spline_trajectory_optimization_b200 import candidates, tracks
    from spline_trajectory_optimization_b200.models.race_track import RaceTrack
    from spline_trajectory_optimization_b200.models.vehicle import Vehicle
    c, l, r = tracks.monza_raw()
    rt = RaceTrack("monza", l, r, c, s=10.0, interval=0.25)
    M = len(rt.center_d)
    assert 23000 < M < 23300
    B = 4
    off = candidates.smooth_offsets(M, B, rt.dist_to_left, rt.dist_to_right, seed=11)
    nrm = rt.left_normals()
    ev = sto.BatchedLineEvaluator(rt.center_d[:, :2], nrm, rt.center_d.ts(), Vehicle(test_vehicle_params()))
    from spline_trajectory_optimization_b200 import _lib
    lib = _lib.load()
    lap, st = ev.lap_times(to_sm(off), B=B)       # default solver: FITPACK's Givens sweep, a 116 k-rotation chain per line
    lap = lap.cpu().numpy()
    assert not st.cpu().numpy().any()
    g = golden("cand_m579_n579")
    ov = O.make_vehicle(*veh_args(g))
    olap, ost = O.lap_batch(rt.center_d[:, 0], rt.center_d[:, 1], nrm[:, 0], nrm[:, 1], off, rt.center_d.ts(),
                            np.zeros(M), ov, n_threads=4, ref_pow=0)
    assert not ost.any() and np.array_equal(lap, olap)
    for solver in ("partitioned", "thomas"):   # 32 lanes share each line's cyclic solve / the one-lane chain of 23 k rows
        with fit_solver(solver):
            if solver == "partitioned":
                assert lib.sto_fit_solver_lanes(M, 2) == 32
            lap_s, st_s = ev.lap_times(to_sm(off[:2]), B=2)
            olap_s, _ = O.lap_batch(rt.center_d[:, 0], rt.center_d[:, 1], nrm[:, 0], nrm[:, 1], off[:2], rt.center_d.ts(),
                                    np.zeros(M), ov, n_threads=2, ref_pow=0)
        assert not st_s.cpu().numpy().any() and np.array_equal(lap_s.cpu().numpy(), olap_s)
        assert np.max(np.abs(olap_s - olap[:2])) < 1e-4       # schedule bifurcations allowed


def test_fit_partitioned_on_device(sto):
    """Default fit solver: lane groups solve the cyclic system together (block elimination + PCR over __shfl_sync).
    Bit-exact against the host emulation for EVERY lane count (the block structure depends on M only, so a line's
    coefficients never depend on the batch it was evaluated in); knots identical to FITPACK; coefficients within 1e-12
    of the Thomas solve / 1e-9 of FITPACK; laps bit-exact against the oracle's sampler + QSS run on these coefficients."""
    import os
    import hostsim_py as H
    from spline_trajectory_optimization_b200 import _lib
    lib = _lib.load()
    for name in ("cand_m579_n579", "cand_m2895_n2895"):
        d = golden(name)
        ev = _evaluator(sto, d)
        B, M = d["offsets"].shape
        hu, hcx, hcy, _ = H.fit_points(d["points"], -1)
        with fit_solver("partitioned"):
            assert lib.sto_fit_solver_lanes(M, B) == 32 and lib.sto_fit_solver_lanes(M, 4096) == 8
            assert lib.sto_fit_solver_lanes(M, 1 << 20) == 1 and lib.sto_fit_solver_lanes(100, 4) == 1
            try:
                for lanes in (32, 16, 8, 4, 2, 1):
                    _lib.check(lib.sto_set_tuning(b"fit_split", lanes))
                    u, cx, cy, st = ev.fit(to_sm(d["offsets"]), B=B)
                    torch.cuda.synchronize()
                    assert not st[:B].any()
                    u, cx, cy = to_cm(u, B), to_cm(cx, B), to_cm(cy, B)
                    assert np.array_equal(u, hu) and np.array_equal(cx, hcx) and np.array_equal(cy, hcy), lanes
            finally:
                _lib.check(lib.sto_set_tuning(b"fit_split", 0))
            lap1, st1 = ev.lap_times(to_sm(d["offsets"]), B=B)
        with fit_solver("thomas"):
            u0, cx0, cy0, _ = ev.fit(to_sm(d["offsets"]), B=B)
            lap0, _ = ev.lap_times(to_sm(d["offsets"]), B=B)
            torch.cuda.synchronize()
        assert np.array_equal(u, to_cm(u0, B))
        assert rel_err(cx, to_cm(cx0, B)) < 1e-12 and rel_err(cy, to_cm(cy0, B)) < 1e-12
        assert rel_err(cx, d["ref_cx"]) < 1e-9 and rel_err(cy, d["ref_cy"]) < 1e-9
        lap1 = lap1[:B].cpu().numpy()
        assert not st1[:B].cpu().numpy().any()
        assert np.max(np.abs(lap1 - lap0[:B].cpu().numpy())) < 1e-7
        assert np.max(np.abs(lap1 - d["ref_lap"])) < 1e-6
        ov = O.make_vehicle(*veh_args(d))
        for b in range(B):
            assert lap1[b] == oracle_lap_from_coefficients(O, u[b], cx[b], cy[b], d["ts"], None, ov)


def test_fit_partitioned_quarter_metre(sto):
    """BASELINE config 5 fit size (M = 23,160): 32 lanes per line for a small batch, fewer as the batch grows - with
    identical bits; equal to the host emulation bit for bit and to the oracle's Thomas solve within 1e-12."""
    import hostsim_py as H
    from spline_trajectory_optimization_b200 import _lib, candidates, tracks
    from spline_trajectory_optimization_b200.models.race_track import RaceTrack
    from spline_trajectory_optimization_b200.models.vehicle import Vehicle
    lib = _lib.load()
    c, l, r = tracks.monza_raw()
    rt = RaceTrack("monza", l, r, c, s=10.0, interval=0.25)
    M = len(rt.center_d)
    nrm = rt.left_normals()
    ev = sto.BatchedLineEvaluator(rt.center_d[:, :2], nrm, rt.center_d.ts(), Vehicle(test_vehicle_params()))
    off = candidates.smooth_offsets(M, 600, rt.dist_to_left, rt.dist_to_right, seed=3)
    hu, hcx, hcy, _ = H.fit_offsets(rt.center_d[:, 0], rt.center_d[:, 1], nrm[:, 0], nrm[:, 1], off[:3], split=-1)
    with fit_solver("partitioned"):
        assert lib.sto_fit_solver_lanes(M, 4) == 32 and lib.sto_fit_solver_lanes(M, 20000) == 4
        assert lib.sto_fit_solver_lanes(M, 40000) == 2
        for B in (1, 3, 70, 600):       # 32, 32, 32 and 16 lanes per line
            u, cx, cy, st = ev.fit(to_sm(off[:B]), B=B)
            torch.cuda.synchronize()
            assert not st[:B].any()
            u, cx, cy = to_cm(u, B), to_cm(cx, B), to_cm(cy, B)
            nb = min(B, 3)
            assert np.array_equal(u[:nb], hu[:nb]) and np.array_equal(cx[:nb], hcx[:nb]) and np.array_equal(cy[:nb], hcy[:nb])
    pts = rt.center_d[None, :, :2] + off[:3, :, None] * nrm[None]
    fu, fcx, fcy, fst = ev.fit(to_sm(off[:3]), B=3)            # default: FITPACK's sweep, 23 k rows
    fcx, fcy = to_cm(fcx, 3), to_cm(fcy, 3)
    for b in range(3):
        t, ocx, ocy = O.fit_periodic_cubic(pts[b])             # oracle default = FITPACK restatement
        assert np.array_equal(hu[b], t[3:-3])
        assert np.array_equal(fcx[b], ocx) and np.array_equal(fcy[b], ocy)
        assert rel_err(hcx[b], ocx) < 1e-12 and rel_err(hcy[b], ocy) < 1e-12


@pytest.mark.parametrize("k,s", [(3, 10.0), (5, 30.0)])
def test_fit_lsq_on_device(sto, k, s):
    """SURVEY.md section 8 f-4 (non-interpolating candidates): least-squares closed splines on the shared knots of a
    smoothing fit.  Device == host build of the same code bit for bit; == FITPACK task = -1 (SciPy, run here as the
    checker) within 1e-12; laps through the shared-knot sampler + QSS equal the reference sequence
    BSplineTrajectory-like spline -> sample_along(ts) -> run_simulation evaluated by the oracle on SciPy's coefficients."""
    import warnings
    import hostsim_py as H
    from scipy.interpolate import splprep
    d = golden("cand_m2895_n2895")
    ev = _evaluator(sto, d)
    B, M = d["offsets"].shape
    pts = d["points"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c0 = np.vstack([pts[0], pts[0][:1]])
        t = splprep([c0[:, 0], c0[:, 1]], s=s, k=k, per=1)[0][0]
        ref = [splprep([np.append(p[:, 0], p[0, 0]), np.append(p[:, 1], p[0, 1])], task=-1, t=t, k=k, per=1) for p in pts]
    u, cx, cy, st = ev.fit_lsq(t, k, offsets_sm=to_sm(d["offsets"]), B=B)
    assert not st[:B].any()
    u, cx, cy = to_cm(u, B), to_cm(cx, B), to_cm(cy, B)
    hu, hcx, hcy, hst = H.fit_lsq(pts, t, k)
    # the host build sums p = centre + o * n identically only when fed the same points: compare through explicit points
    u2, cx2, cy2, st2 = ev.fit_lsq(t, k, points_sm=(to_sm(pts[:, :, 0]), to_sm(pts[:, :, 1])), B=B)
    assert np.array_equal(to_cm(u2, B), hu) and np.array_equal(to_cm(cx2, B), hcx) and np.array_equal(to_cm(cy2, B), hcy)
    for b in range(B):
        (tt, (rx, ry), kk), ru = ref[b]
        assert np.max(np.abs(u[b] - ru)) < 1e-15
        assert np.max(np.abs(cx[b] - rx)) < 1e-12 * np.max(np.abs(rx)) and np.max(np.abs(cy[b] - ry)) < 1e-12 * np.max(np.abs(ry))
    lap, lst = ev.lap_times_lsq(t, k, to_sm(d["offsets"]), B=B)
    lap = lap.cpu().numpy()
    assert not lst.cpu().numpy().any()
    ov = O.make_vehicle(*veh_args(d))
    for b in range(B):
        (tt, (rx, ry), kk), ru = ref[b]
        X, Y, YAW, R = O.sample(t, rx, ry, k, d["ts"], 0)
        olap = O.qss(X, Y, R, np.zeros(len(d["ts"])), ov, 0)["lap"]
        assert abs(lap[b] - olap) < 1e-6                      # BASELINE: lap within 1e-6 s
        X, Y, YAW, R = O.sample(t, cx[b], cy[b], k, d["ts"], 0)
        assert lap[b] == O.qss(X, Y, R, np.zeros(len(d["ts"])), ov, 0)["lap"]   # identical coefficients: bit-exact
    # a smoothed line is not the interpolating one: the laps differ from the s = 0 path by a visible amount
    lap0, _ = ev.lap_times(to_sm(d["offsets"]), B=B)
    assert np.max(np.abs(lap0[:B].cpu().numpy() - lap)) > 1e-3
    with pytest.raises(ValueError):
        ev.fit_lsq(t[:12], k, offsets_sm=to_sm(d["offsets"]), B=B)


def test_status_flags_and_chunked_host_path(sto):
    """Data-dependent failures are per-candidate status bits, never a batch abort; the host entry point chunks a batch
    to a byte budget without changing a bit."""
    from spline_trajectory_optimization_b200 import _lib
    d = golden("cand_m579_n579")
    veh = _lib.make_vehicle(*veh_args(d))
    N = len(d["ts"])
    X, Y, R = d["ref_X"][:3].copy(), d["ref_Y"][:3].copy(), d["ref_CURVATURE"][:3].copy()
    R[1, 100] = 0.0            # zero turn radius -> zero speed at that sample -> the reference divides by zero (:164-165)
    X[2, 50] = np.nan          # NaN position -> NaN chord -> NaN segment time
    for impl in ("plain", "memo"):
        res = sto.run_qss(to_sm(X), to_sm(Y), to_sm(R), veh, B=3, impl=sto.IMPL[impl])
        st = res["status"][:3].cpu().numpy()
        lap = res["lap"][:3].cpu().numpy()
        assert st[0] == 0 and abs(lap[0] - d["ref_lap"][0]) < 1e-9          # the healthy neighbour is untouched
        assert st[1] & _lib.CAND_ZERO_SPEED and np.isnan(lap[1])
        assert st[2] != 0 and np.isnan(lap[2])
    # the single-trajectory mirror raises what the reference raises
    from spline_traj_optm.models.trajectory import Trajectory
    from spline_traj_optm.models.vehicle import Vehicle
    from spline_traj_optm.simulator.simulator import Simulator
    traj = Trajectory(N)
    traj[:, Trajectory.X], traj[:, Trajectory.Y], traj[:, Trajectory.CURVATURE] = X[1], Y[1], R[1]
    with pytest.raises(FloatingPointError):
        Simulator(Vehicle(test_vehicle_params())).run_simulation(traj)
    # chunking: a budget that forces several passes over the batch gives the same laps
    ev = _evaluator(sto, d)
    lap_one, st_one = ev.lap_times_host(np.tile(d["offsets"], (12, 1)))
    need32 = ev.lib.sto_lap_workspace_bytes(ev.M, ev.N, 32, ev.impl)
    lap_chunks, st_chunks = ev.lap_times_host(np.tile(d["offsets"], (12, 1)), max_work_bytes=int(need32 * 1.5) + (4 << 20))
    assert np.array_equal(lap_one, lap_chunks) and not st_one.any() and not st_chunks.any()
    assert np.array_equal(lap_one[:8], lap_one[88:])


@pytest.mark.parametrize("kernel", [2, 3, 4])
def test_failed_line_shares_a_warp_with_healthy_lines(sto, kernel):
    """Zero-radius samples (zero initial speed -> the reference's division by zero, status ZERO_SPEED) at different places of
    some lines of a batch that is forced onto 4 lines x 8 lanes per warp: the failing lines stop with their status bit and a
    NaN lap at whatever phase of the round loop they fail in, the launch terminates, and the healthy lines of the same warps
    equal the oracle bit for bit (laps, speeds, step and iteration counts)."""
    from spline_trajectory_optimization_b200 import _lib
    lib = _lib.load()
    d = golden("cand_m579_n579")
    veh, ov = _lib.make_vehicle(*veh_args(d)), O.make_vehicle(*veh_args(d))
    B = 16
    X, Y, R = (np.tile(d[k][:8], (2, 1)).copy() for k in ("ref_X", "ref_Y", "ref_CURVATURE"))
    R *= (1.0 + 0.01 * np.arange(B))[:, None]
    bad = {1: [100], 6: [3, 250, 251, 400], 9: [578], 15: [0, 1, 2]}
    for c, where in bad.items():
        R[c, where] = 0.0
    try:
        _lib.check(lib.sto_set_tuning(b"qss_kernel", kernel))
        _lib.check(lib.sto_set_tuning(b"qss_group", 8))
        res = sto.run_qss(to_sm(X), to_sm(Y), to_sm(R), veh, B=B, impl=sto.IMPL["memo"])
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.sto_set_tuning(b"qss_kernel", 0))
        _lib.check(lib.sto_set_tuning(b"qss_group", 0))
    st, lap = res["status"][:B].cpu().numpy(), res["lap"][:B].cpu().numpy()
    for c in range(B):
        if c in bad:
            assert st[c] & _lib.CAND_ZERO_SPEED and np.isnan(lap[c]), c
            continue
        o = O.qss(X[c], Y[c], R[c], np.zeros(X.shape[1]), ov, 0)
        assert st[c] == 0 and lap[c] == o["lap"], c
        assert np.array_equal(res["speed"][:, c].cpu().numpy(), o["v"]), c
        assert float(res["summary"][6, c]) == o["steps"] and float(res["summary"][7, c]) == o["iters"] + 1, c


@pytest.mark.parametrize("case", ["nan_radius", "nan_position", "all_nan", "nan_mixed"])
def test_qss_nan_inputs_equal_oracle(sto, case):
    """NaN inside a live line: the reference's spawn test (simulator.py:239, `g > max_curve_speed or g < min_state_speed`)
    is false for NaN, so such a front only stops.  Both kernels must EQUAL the oracle - status, front-step count, outer
    iterations, speeds and accelerations (NaN pattern included) - with the healthy neighbours in the batch untouched."""
    from helpers import nan_cases
    from spline_trajectory_optimization_b200 import _lib
    d = golden("sim_s10k3_i2")
    veh, ov = _lib.make_vehicle(*veh_args(d)), O.make_vehicle(*veh_args(d))
    x, y, r, sb = nan_cases()[case]
    from helpers import synthetic_closed_track
    xh, yh, rh, _ = synthetic_closed_track(3, 300)             # the same line without the NaN, as lanes 0 and 2
    X, Y, R = np.stack([xh, x, xh]), np.stack([yh, y, yh]), np.stack([rh, r, rh])
    o, oh = O.qss(x, y, r, sb, ov, 0), O.qss(xh, yh, rh, sb, ov, 0)
    assert o["status"] == 2 and oh["status"] == 0
    for impl in ("plain", "memo"):
        res = sto.run_qss(to_sm(X), to_sm(Y), to_sm(R), veh, B=3, sin_bank=sb, impl=sto.IMPL[impl])
        torch.cuda.synchronize()
        st = res["status"][:3].cpu().numpy()
        assert list(st) == [0, _lib.CAND_NAN, 0], (impl, st)
        assert float(res["summary"][6, 1]) == o["steps"] and float(res["summary"][7, 1]) == o["iters"] + 1, impl
        assert np.array_equal(res["speed"][:, 1].cpu().numpy(), o["v"], equal_nan=True), impl
        assert np.array_equal(res["lon_acc"][:, 1].cpu().numpy(), o["a"], equal_nan=True), impl
        assert np.isnan(float(res["lap"][1]))
        for c in (0, 2):
            assert np.array_equal(res["speed"][:, c].cpu().numpy(), oh["v"]) and float(res["lap"][c]) == oh["lap"]
            assert float(res["summary"][6, c]) == oh["steps"]


def test_fused_path_nan_offset_terminates(sto):
    """One NaN offset makes every coefficient of that line NaN (the fit is periodic): the fit flags it, the QSS must not
    spin on it (round 1 re-spawned every row until the 64 N iteration cap: minutes for one line at N = 2895) and the
    other lines of the batch keep their bits.  Status = DEGENERATE_FIT | NAN, lap NaN, the oracle agrees."""
    import time
    from spline_trajectory_optimization_b200 import _lib
    d = golden("cand_m2895_n2895")
    ov = O.make_vehicle(*veh_args(d))
    nrm = np.stack([-np.sin(d["centre_yaw"]), np.cos(d["centre_yaw"])], axis=1)
    off = d["offsets"].copy()
    B = off.shape[0]
    off[1, 1000] = np.nan
    olap, ost = O.lap_batch(d["centre_x"], d["centre_y"], nrm[:, 0], nrm[:, 1], off, d["ts"], np.zeros(len(d["ts"])),
                            ov, n_threads=2, ref_pow=0)
    assert ost[1] != 0 and np.isnan(olap[1])
    for impl in ("memo", "plain"):
        ev = _evaluator(sto, d, impl=impl)
        ev.lap_times(to_sm(d["offsets"]), B=B)                 # warm-up (module load, workspace)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lap, st = ev.lap_times(to_sm(off), B=B)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        lap, st = lap.cpu().numpy(), st.cpu().numpy()
        assert st[1] == (_lib.CAND_DEGENERATE_FIT | _lib.CAND_NAN) and np.isnan(lap[1]), (impl, st)
        keep = np.arange(B) != 1
        assert not st[keep].any() and np.array_equal(lap[keep], olap[keep]), impl
        assert dt < (5.0 if impl == "memo" else 20.0), (impl, dt)


def test_host_entry_is_thread_safe_and_overlaps(sto):
    """sto_lap_time_host_f64 from several host threads at once (the library keeps two contexts per device: two calls
    overlap on their own streams, further callers queue): every call returns exactly what a lone call returns."""
    import threading
    d = golden("cand_m579_n579")
    ev = _evaluator(sto, d)
    batches = [np.tile(d["offsets"], (k + 1, 1)) * (1.0 - 0.01 * k) for k in range(6)]
    want = [ev.lap_times_host(b) for b in batches]
    got = [None] * len(batches)

    def worker(k):
        for _ in range(3):
            got[k] = ev.lap_times_host(batches[k])

    th = [threading.Thread(target=worker, args=(k,)) for k in range(len(batches))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for k in range(len(batches)):
        assert np.array_equal(got[k][0], want[k][0]) and not got[k][1].any(), k


def test_device_argmin_pair_kernels(sto):
    """sharding.DeviceArgmin (sto_argmin_pair_f64 + sto_argmin_pairs_f64; one rank here - the gathered-pairs kernel is
    also fed four hand-made rank pairs): NaN laps and failed candidates ignored, ties to the lowest global index, an
    all-invalid shard gives (NaN, -1)."""
    import ctypes as C
    from spline_trajectory_optimization_b200 import _lib, sharding
    lib = _lib.load()
    dev = torch.device("cuda")
    red = sharding.DeviceArgmin(dev)
    lap = torch.tensor([105.5, float("nan"), 104.25, 104.25, 103.0, 110.0], dtype=torch.float64, device=dev)
    st = torch.tensor([0, 0, 0, 0, 4, 0], dtype=torch.int32, device=dev)        # candidate 4 failed: not eligible
    best, idx = red(lap, st, 1000)
    assert float(best[0]) == 104.25 and int(idx[0]) == 1002
    best, idx = red(torch.full((5,), float("nan"), dtype=torch.float64, device=dev), None, 7)
    assert np.isnan(float(best[0])) and int(idx[0]) == -1
    pairs = torch.tensor([107.0, 12.0, float("nan"), -1.0, 106.5, 4100.0, 106.5, 300.0], dtype=torch.float64, device=dev)
    ob, oi = torch.empty(1, dtype=torch.float64, device=dev), torch.empty(1, dtype=torch.int64, device=dev)
    _lib.check(lib.sto_argmin_pairs_f64(C.c_void_p(pairs.data_ptr()), 4, C.c_void_p(ob.data_ptr()),
                                        C.c_void_p(oi.data_ptr()), None))
    torch.cuda.synchronize()
    assert float(ob[0]) == 106.5 and int(oi[0]) == 300


def test_control_point_variants_batched(sto):
    """§8 f-2 (optimiser-loop batching): the reference's edit -> wrap -> sample_along(ts) -> run_simulation sequence for
    six control-point variants of the s=30,k=5 Monza line, scored in one launch."""
    from spline_trajectory_optimization_b200 import tracks
    from spline_trajectory_optimization_b200.models.trajectory import BSplineTrajectory
    from spline_trajectory_optimization_b200.models.vehicle import Vehicle
    from spline_trajectory_optimization_b200.optimization.batched import score_control_point_variants
    d = golden("ctrl_s30k5_n578")
    spl = BSplineTrajectory(tracks.monza_raw()[0], 30.0, 5)
    edits = [{int(i): tuple(xy) for i, xy in zip(idx, xys) if i >= 0} for idx, xys in zip(d["edit_idx"], d["edit_xy"])]
    lap, st = score_control_point_variants(spl, edits, d["ts"], Vehicle(test_vehicle_params()))
    assert not st.any()
    assert np.max(np.abs(lap - d["ref_lap"])) < 1e-9          # identical inputs (same coefficients): 1e-9 s
    S = sto.sample_splines(d["t"], int(d["k"]), to_sm(d["ref_cx"]), to_sm(d["ref_cy"]), d["ts"], B=6)
    assert np.array_equal(to_cm(S["x"], 6), d["ref_X"]) and np.array_equal(to_cm(S["y"], 6), d["ref_Y"])
    assert rel_err(to_cm(S["radius"], 6), d["ref_CURVATURE"]) < 1e-15


def test_fill_bounds_on_device(sto):
    """§8 f-1: Trajectory.fill_bounds on the GPU == its NumPy restatement (bit for bit) on the Monza track."""
    from spline_trajectory_optimization_b200 import tracks
    from spline_trajectory_optimization_b200.models.race_track import RaceTrack
    from spline_trajectory_optimization_b200.models.trajectory import Trajectory
    c, l, r = tracks.monza_raw()
    rt = RaceTrack("monza", l, r, c, s=10.0, interval=5.0)       # constructor already used the device path
    host = rt.center_d.copy()
    host[:, Trajectory.LEFT_BOUND_X:Trajectory.RIGHT_BOUND_Y + 1] = 0.0
    host.fill_bounds(rt.left_r, rt.right_r, max_dist=100.0, device=False)
    cols = slice(Trajectory.LEFT_BOUND_X, Trajectory.RIGHT_BOUND_Y + 1)
    assert np.array_equal(host[:, cols], rt.center_d[:, cols])
    assert 3.0 < rt.dist_to_left.min() and rt.dist_to_left.max() < 7.5 and 3.0 < rt.dist_to_right.min()
    # a point with no boundary within max_dist keeps itself as its bound (reference :118-127)
    far = Trajectory(1)
    far[0, Trajectory.X], far[0, Trajectory.Y] = 1e5, 1e5
    far.fill_bounds(rt.left_r, rt.right_r, max_dist=100.0)
    assert far[0, Trajectory.LEFT_BOUND_X] == 1e5 and far[0, Trajectory.RIGHT_BOUND_Y] == 1e5


@pytest.mark.parametrize("closed", [False, True])
def test_fill_bounds_reference_semantics_on_device(sto, closed):
    """The hand-derived fill_bounds cases of tests/test_track_io.py (reference models/trajectory.py:83-141: two-sided
    200 m segment, nearest Point wins whichever side it is on, no hit -> the point itself, ring with / without the
    repeated closing vertex) through sto_fill_bounds_f64, and device == NumPy path bit for bit."""
    from test_track_io import _closed, fill_bounds_cases
    from spline_trajectory_optimization_b200.models.trajectory import Trajectory
    for P, yaw, lring, rring, eleft, eright, fl, fr in fill_bounds_cases():
        lr, rr = (_closed(lring), _closed(rring)) if closed else (lring, rring)
        out = {}
        for device in (True, False):
            t = Trajectory(1)
            t.points[0, [Trajectory.X, Trajectory.Y]] = P
            t.points[0, Trajectory.YAW] = yaw
            t.fill_bounds(lr, rr, max_dist=100.0, device=device)
            out[device] = t.points[0, 9:13].copy()
        assert np.array_equal(out[True], out[False]), (P, out)
        assert np.allclose(out[True][:2], eleft, atol=1e-12) and np.allclose(out[True][2:], eright, atol=1e-12), (P, out)


def test_fast_fp64_division_and_sqrt_selftest(sto):
    """The FITPACK solver's rotation chain runs its divisions and square roots as flagged, branch-free copies of nvcc's
    fast paths (csrc/sto_common.cuh).  On 3 x 2^27 operations over every kind of operand: not one result differs from the
    plain operators while the flag is down; the flag is up only for operands outside the fast path's range."""
    import ctypes
    from spline_trajectory_optimization_b200 import _lib
    lib = _lib.load()
    for seed in (1, 2, 3):
        tested, bad, flagged = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_longlong()
        _lib.check(lib.sto_selftest_fp64(seed, ctypes.byref(tested), ctypes.byref(bad), ctypes.byref(flagged)))
        assert tested.value == 1 << 27 and bad.value == 0
        assert 0 < flagged.value < tested.value // 2      # random bit patterns hit the guards; in-range operands do not
