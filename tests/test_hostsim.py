"""CPU: the per-candidate device functions (csrc/*.cuh) compiled for the host, against the oracle.  Exercises the
kernel logic (implicit row positions, re-spawned row table, ring windows, edge memoisation) without a GPU.  The
oracle runs in ref_pow=0 mode (plain products), the arithmetic the kernels use: everything must match BIT FOR BIT."""
import numpy as np
import pytest

import hostsim_py as H
import oracle_py as O
from helpers import CAND_CASES, SIM_CASES, golden, rel_err, veh_args


@pytest.fixture(autouse=True)
def _default_oracle_solver():
    yield
    O.set_fit_solver("fitpack")


@pytest.mark.parametrize("name", CAND_CASES[:2])
@pytest.mark.parametrize("solver", ["fitpack", "thomas", "blocks"])
def test_fit_and_eval_bit_exact(name, solver):
    """Host build of the device fit (all three solvers) == the oracle's restatement of the same solver, bit for bit;
    the FITPACK solver == the reference's own coefficients (golden, from the unmodified reference) bit for bit."""
    d = golden(name)
    O.set_fit_solver(solver)
    u, cx, cy, st = H.fit_points(d["points"], {"thomas": 1, "blocks": -1, "fitpack": -108}[solver])
    assert not st.any()
    if solver == "fitpack":
        assert np.array_equal(cx, d["ref_cx"]) and np.array_equal(cy, d["ref_cy"])
        assert np.array_equal(H.fit_points(d["points"], -101)[1], cx)        # any split of the row phases
    E = H.evaluate(u, cx, cy, d["ts"])
    for b in range(d["points"].shape[0]):
        t, ocx, ocy = O.fit_periodic_cubic(d["points"][b])
        assert np.array_equal(u[b], t[3:-3]) and np.array_equal(cx[b], ocx) and np.array_equal(cy[b], ocy)
        X, Y, YAW, R = O.sample(t, ocx, ocy, 3, d["ts"], 0)
        assert np.array_equal(E["x"][b], X) and np.array_equal(E["y"][b], Y) and np.array_equal(E["radius"][b], R)
        assert np.max(np.abs(E["yaw"][b] - YAW)) < 1e-15


def test_fit_from_offsets_equals_fit_from_points():
    d = golden("cand_m579_n579")
    nx, ny = -np.sin(d["centre_yaw"]), np.cos(d["centre_yaw"])
    u1, cx1, cy1, _ = H.fit_offsets(d["centre_x"], d["centre_y"], nx, ny, d["offsets"])
    u2, cx2, cy2, _ = H.fit_points(d["points"])
    assert np.array_equal(u1, u2) and np.array_equal(cx1, cx2) and np.array_equal(cy1, cy2)


@pytest.mark.parametrize("name", SIM_CASES)
@pytest.mark.parametrize("impl", [0, 1])
def test_qss_both_schedules_bit_exact(name, impl):
    d = golden(name)
    ov, hv = O.make_vehicle(*veh_args(d)), H.make_vehicle(*veh_args(d))
    sb = np.sin(d["in_BANK"])
    o = O.qss(d["in_X"], d["in_Y"], d["in_CURVATURE"], sb, ov, 0)
    r = H.qss(impl, d["in_X"][None], d["in_Y"][None], d["in_CURVATURE"][None], sb, hv, owner=(impl == 0))
    assert r["status"][0] == 0
    for k, ok in (("v", "v"), ("a", "a"), ("lat", "lat"), ("time", "time")):
        assert np.array_equal(r[k][0], o[ok]), k
    assert r["lap"][0] == o["lap"]
    assert r["summary"][0, 6] == o["steps"] and r["summary"][0, 7] == o["iters"] + 1
    if impl == 0:
        assert np.array_equal(r["owner"][0], o["flag"])


@pytest.mark.parametrize("impl", [0, 1])
def test_qss_synthetic_tables(impl):
    d = golden("sim_synthetic_tables")
    ov, hv = O.make_vehicle(*veh_args(d)), H.make_vehicle(*veh_args(d))
    for tag in ("n8", "n64", "n257"):
        sb = np.sin(d[tag + "_in_BANK"])
        o = O.qss(d[tag + "_in_X"], d[tag + "_in_Y"], d[tag + "_in_CURVATURE"], sb, ov, 0)
        r = H.qss(impl, d[tag + "_in_X"][None], d[tag + "_in_Y"][None], d[tag + "_in_CURVATURE"][None], sb, hv)
        assert np.array_equal(r["v"][0], o["v"]) and np.array_equal(r["time"][0], o["time"])


def test_generic_spline_eval_k5():
    d = golden("sim_s30k5_i5")
    x, y, yaw, rad = H.eval_spline(d["spl_t"], d["spl_cx"], d["spl_cy"], 5, d["ts"])
    assert np.array_equal(x, d["in_X"]) and np.array_equal(y, d["in_Y"])
    assert np.max(np.abs(rad - d["in_CURVATURE"]) / d["in_CURVATURE"]) < 1e-15


@pytest.mark.parametrize("group", [2, 4, 8])
@pytest.mark.parametrize("name", ["sim_s10k3_i2", "sim_s30k5_i3", "sim_oval_bank12"])
def test_qss_lane_group_emulation_bit_exact(name, group):
    """Lane groups: the backward sub-pass evaluates up to G consecutive dirty fronts of a word against the state as it
    stood and commits them in row order (memo_bwd_rows_group).  Host emulation of the G lanes must reproduce the
    one-at-a-time schedule bit for bit, including the front-step count."""
    d = golden(name)
    ov, hv = O.make_vehicle(*veh_args(d)), H.make_vehicle(*veh_args(d))
    sb = np.sin(d["in_BANK"])
    o = O.qss(d["in_X"], d["in_Y"], d["in_CURVATURE"], sb, ov, 0)
    r = H.qss(100 + group, d["in_X"][None], d["in_Y"][None], d["in_CURVATURE"][None], sb, hv)
    assert r["status"][0] == 0
    for k in ("v", "a", "lat", "time"):
        assert np.array_equal(r[k][0], o[k]), k
    assert r["lap"][0] == o["lap"] and r["summary"][0, 6] == o["steps"]
    assert r["summary"][0, 7] == o["iters"] + 1     # (a trailing tombstone in the backward list must not cost an iteration)


@pytest.mark.parametrize("n", [128, 129, 191, 192, 193, 256, 257, 321, 1024, 1025])
def test_qss_memo_word_boundaries(n):
    """Track sizes around multiples of 64 (the bit planes' word size): the dense iteration-0 forward sweep collects plane
    bits per word and leaves the seam row to the general walker; the list walkers own entries modulo the group size.
    One lane per candidate and emulated lane groups against the oracle, bit for bit, including the step count."""
    d = golden("sim_s10k3_i2")
    ov, hv = O.make_vehicle(*veh_args(d)), H.make_vehicle(*veh_args(d))
    idx = np.linspace(0, len(d["in_X"]) - 1, n, endpoint=False).astype(int)
    x, y, r = d["in_X"][idx].copy(), d["in_Y"][idx].copy(), d["in_CURVATURE"][idx].copy()
    sb = np.zeros(n)
    o = O.qss(x, y, r, sb, ov, 0)
    for impl in (1, 104, 108):
        h = H.qss(impl, x[None], y[None], r[None], sb, hv)
        assert h["status"][0] == 0
        for k in ("v", "a", "lat", "time"):
            assert np.array_equal(h[k][0], o[k]), (impl, k)
        assert h["lap"][0] == o["lap"] and h["summary"][0, 6] == o["steps"]


@pytest.mark.parametrize("seed", range(8))
def test_qss_memo_random_tracks(seed):
    """Random synthetic tracks (hairpins, banked stretches, 128-700 samples): the memoised schedule with one lane per line
    and with emulated lane groups of 2 / 4 / 8 against the oracle, bit for bit, including the step count."""
    from helpers import synthetic_closed_track
    d = golden("sim_s10k3_i2")
    ov, hv = O.make_vehicle(*veh_args(d)), H.make_vehicle(*veh_args(d))
    n = int(np.random.default_rng(100 + seed).integers(128, 700))
    x, y, r, sb = synthetic_closed_track(seed, n)
    o = O.qss(x, y, r, sb, ov, 0)
    for impl in (1, 102, 104, 108):
        h = H.qss(impl, x[None], y[None], r[None], sb, hv)
        assert h["status"][0] == 0
        for k in ("v", "a", "lat", "time"):
            assert np.array_equal(h[k][0], o[k]), (impl, k)
        assert h["lap"][0] == o["lap"] and h["summary"][0, 6] == o["steps"] and h["summary"][0, 7] == o["iters"] + 1, impl


def test_forward_list_batch_prototype_is_exact():
    """Next-round prototype (host only, memo_spawned_fwd_batch): the forward re-spawned list evaluated up to G fronts at a
    time, resolved from start-of-segment memo bits plus the members' outcomes, cut where a member writes a later entry's
    sample.  Must reproduce the one-at-a-time walk bit for bit (state, list, step count); reports evaluations per batch."""
    import ctypes as C
    from helpers import synthetic_closed_track
    L = H.lib()
    d = golden("sim_s10k3_i2")
    ov, hv = O.make_vehicle(*veh_args(d)), H.make_vehicle(*veh_args(d))
    cases = [(d["in_X"], d["in_Y"], d["in_CURVATURE"], np.sin(d["in_BANK"]))]
    for seed in range(6):
        n = int(np.random.default_rng(5000 + seed).integers(128, 1500))
        cases.append(synthetic_closed_track(7000 + seed, n))
    out = (C.c_longlong * 4)()
    try:
        L.hostsim_fwd_batch(1)
        L.hostsim_fwd_batch_counters(out, 1)
        for x, y, r, sb in cases:
            o = O.qss(x, y, r, sb, ov, 0)
            for impl in (102, 104, 108):
                h = H.qss(impl, x[None], y[None], r[None], sb, hv)
                assert h["status"][0] == 0
                for k in ("v", "a", "lat", "time"):
                    assert np.array_equal(h[k][0], o[k]), (impl, k)
                assert h["lap"][0] == o["lap"] and h["summary"][0, 6] == o["steps"]
        L.hostsim_fwd_batch_counters(out, 1)
        assert out[1] > 0 and out[0] > 0        # members were committed in batches
    finally:
        L.hostsim_fwd_batch(0)


@pytest.mark.parametrize("split", [2, 7, 16])
def test_eval_sample_range_split(split):
    """Splitting a candidate's samples over lanes (eval_range) gives the same bits as one walk over all of them,
    including the chords across slice boundaries and the closing chord."""
    d = golden("cand_m579_n1158")
    u, cx, cy, _ = H.fit_points(d["points"])
    whole = H.evaluate(u, cx, cy, d["ts"])
    parts = H.evaluate(u, cx, cy, d["ts"], split=split)
    for k in whole:
        assert np.array_equal(whole[k], parts[k]), k


@pytest.mark.parametrize("split", [2, 4, 5])
def test_fit_lane_group_phases(split):
    """The fit split into lane-group phases (row work interleaved over lanes, three Thomas recurrences side by side)
    produces the same bits as the one-lane fit."""
    d = golden("cand_m579_n579")
    u1, cx1, cy1, _ = H.fit_points(d["points"])
    u2, cx2, cy2, _ = H.fit_points(d["points"], split=split)
    assert np.array_equal(u1, u2) and np.array_equal(cx1, cx2) and np.array_equal(cy1, cy2)


def test_spline_batch_shared_knots():
    """Coefficient sets on shared knots (the optimiser-loop shape): samples bit-identical to the reference's
    sample_along(ts=...) after set_control_point edits, chords consistent with the per-candidate evaluator."""
    d = golden("ctrl_s30k5_n578")
    E = H.eval_spline_batch(d["t"], int(d["k"]), d["ref_cx"], d["ref_cy"], d["ts"])
    assert np.array_equal(E["x"], d["ref_X"]) and np.array_equal(E["y"], d["ref_Y"])
    assert np.max(np.abs(E["radius"] - d["ref_CURVATURE"]) / d["ref_CURVATURE"]) < 1e-15
    nxt = np.roll(E["x"], -1, axis=1), np.roll(E["y"], -1, axis=1)
    assert np.array_equal(E["chord_qss"], np.sqrt((E["x"] - nxt[0]) ** 2 + (E["y"] - nxt[1]) ** 2))
    hv, ov = H.make_vehicle(*veh_args(d)), O.make_vehicle(*veh_args(d))
    r = H.qss(1, E["x"], E["y"], E["radius"], None, hv)
    for b in range(E["x"].shape[0]):
        o = O.qss(E["x"][b], E["y"][b], E["radius"][b], np.zeros(E["x"].shape[1]), ov, 0)
        assert np.array_equal(r["v"][b], o["v"]) and r["lap"][b] == o["lap"]
        assert abs(r["lap"][b] - d["ref_lap"][b]) < 1e-9


def test_control_point_variants_match_reference_edits():
    from spline_trajectory_optimization_b200.optimization.batched import control_point_variants
    from spline_trajectory_optimization_b200.models.trajectory import BSplineTrajectory
    from spline_trajectory_optimization_b200 import tracks
    d = golden("ctrl_s30k5_n578")
    spl = BSplineTrajectory(tracks.monza_raw()[0], 30.0, 5)
    assert np.array_equal(spl._spl_x.c, d["base_cx"])
    edits = [{int(i): tuple(xy) for i, xy in zip(idx, xys) if i >= 0} for idx, xys in zip(d["edit_idx"], d["edit_xy"])]
    cx, cy = control_point_variants(spl, edits)
    assert np.array_equal(cx, d["ref_cx"]) and np.array_equal(cy, d["ref_cy"])


@pytest.mark.parametrize("name", ["sim_s30k5_i10", "sim_s10k3_i2", "sim_s0k3_i10"])
def test_arc_sections_gauss_legendre(name):
    """DIST_TO_SF_BWD from fixed-order Gauss-Legendre per knot span vs the reference's adaptive quad (its own tolerance
    is 1.5e-8 relative): <= 1e-9 relative."""
    d = golden(name)
    sec = H.arc_sections(d["spl_t"], d["spl_cx"], d["spl_cy"], int(d["spl_k"]), d["ts"])
    bwd = np.cumsum(sec)
    assert bwd[0] == 0.0
    assert np.max(np.abs(bwd[1:] - d["in_DIST_BWD"][1:]) / d["in_DIST_BWD"][1:]) < 1e-9


@pytest.mark.parametrize("name", CAND_CASES)
def test_fit_partitioned_solver(name):
    """Default fit solver (sto_fit.cuh, partitioned cyclic solve): 32 blocks of rows eliminated independently, the
    interface system solved by parallel cyclic reduction.  Same knots bit for bit; coefficients equal to the one-lane
    Thomas solve up to rounding and within BASELINE's 1e-9 of FITPACK; independent of how the row work is split."""
    d = golden(name)
    u0, cx0, cy0, _ = H.fit_points(d["points"], 1)
    ref = None
    assert rel_err(cx0, d["ref_cx"]) < 1e-9
    for lanes in (1, 4, 32):
        u, cx, cy, st = H.fit_points(d["points"], -lanes)
        assert not st.any() and np.array_equal(u, u0)
        assert rel_err(cx, cx0) < 1e-12 and rel_err(cy, cy0) < 1e-12
        assert rel_err(cx, d["ref_cx"]) < 1e-9 and rel_err(cy, d["ref_cy"]) < 1e-9
        assert np.array_equal(cx[:, -3:], cx[:, :3]) and np.array_equal(cy[:, -3:], cy[:, :3])   # periodic wrap
        ref = (cx, cy) if ref is None else ref
        assert np.array_equal(cx, ref[0]) and np.array_equal(cy, ref[1])


def test_fit_partitioned_ragged_and_tiny():
    """Blocks of unequal length, the shortest blocks (8 rows at M = 256), single-block tracks (M < 256) and strongly
    non-uniform spacing."""
    rng = np.random.default_rng(5)
    for M in (4, 7, 16, 64, 255, 256, 257, 1000, 4099):
        th = np.sort(rng.uniform(0, 2 * np.pi, M))
        th[1::3] = th[::3][:len(th[1::3])] + 1e-3 * rng.uniform(0.1, 1.0, len(th[1::3]))   # clustered points
        th = np.sort(th)
        pts = np.stack([np.cos(th) * (1 + 0.1 * rng.standard_normal(M)), 0.7 * np.sin(th)], -1)[None]
        u0, cx0, cy0, st0 = H.fit_points(pts, 1)
        assert not st0.any()
        scale = np.abs(cx0).max()
        u, cx, cy, st = H.fit_points(pts, -1)
        assert not st.any() and np.array_equal(u, u0)
        assert np.max(np.abs(cx - cx0)) < 1e-10 * scale and np.max(np.abs(cy - cy0)) < 1e-10 * scale, M


@pytest.mark.parametrize("k,s", [(3, 10.0), (5, 30.0), (1, 50.0), (4, 20.0)])
def test_fit_lsq_matches_fitpack_fixed_knots(k, s):
    """least-squares closed spline on fixed knots (sto_fit_lsq.cuh) == FITPACK task = -1
    (scipy.interpolate.splprep(task=-1, t=knots, per=1)) on the knots of a smoothing fit, every degree; 1e-9 is the
    BASELINE tolerance for coefficients, 1e-12 is what the bordered Cholesky solve delivers."""
    import warnings
    from scipy.interpolate import splprep
    d = golden("cand_m2895_n2895")
    pts = d["points"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c0 = np.vstack([pts[0], pts[0][:1]])
        t = splprep([c0[:, 0], c0[:, 1]], s=s, k=k, per=1)[0][0]
        u, cx, cy, st = H.fit_lsq(pts, t, k)
        assert not st.any() and cx.shape[1] == len(t) - k - 1
        for b in range(len(pts)):
            c = np.vstack([pts[b], pts[b][:1]])
            (tt, (rx, ry), kk), ru = splprep([c[:, 0], c[:, 1]], task=-1, t=t, k=k, per=1)
            assert np.array_equal(u[b], ru)                                   # chord-length parameter: bit-identical
            assert np.max(np.abs(cx[b] - rx)) < 1e-12 * np.max(np.abs(rx))
            assert np.max(np.abs(cy[b] - ry)) < 1e-12 * np.max(np.abs(ry))
            assert np.array_equal(cx[b, -k:], cx[b, :k]) and np.array_equal(cy[b, -k:], cy[b, :k])


def test_fit_lsq_flags_empty_knot_interval():
    """A knot interval without data: FITPACK returns ier = 10; the candidate is flagged, its neighbours are not."""
    d = golden("cand_m579_n579")
    pts = d["points"][:2].copy()
    k = 3
    inner = np.linspace(0.0, 1.0, 41)
    inner = np.sort(np.concatenate([inner, [0.50001, 0.50002, 0.50003, 0.50004, 0.50005]]))   # five knots between two data points
    per = 1.0
    t = np.concatenate([inner[-k - 1:-1] - per, inner, inner[1:k + 1] + per])
    u, cx, cy, st = H.fit_lsq(pts, t, k)
    assert st.all() and np.isnan(cx).all()
    inner = np.linspace(0.0, 1.0, 41)
    t = np.concatenate([inner[-k - 1:-1] - per, inner, inner[1:k + 1] + per])
    u, cx, cy, st = H.fit_lsq(pts, t, k)
    assert not st.any() and np.isfinite(cx).all()


@pytest.mark.parametrize("M,scale", [(700, 1.0), (1500, 1e-3), (4000, 1e3), (5999, 1.0)])
def test_fit_fitpack_long_lines_decayed_pivots(M, scale):
    """Long closed lines: the fill-in of the two periodic rows decays by ~0.27 per band row, is below 2^-30 of the diagonal
    from row ~16 on (fpgivs in its collapsed form, fp_givs_decayed) and denormal from row ~510 on for the rest of the
    sweep.  Coefficients must still equal SciPy's bit for bit, at any coordinate scale."""
    import warnings
    from scipy.interpolate import splprep
    rng = np.random.default_rng(M)
    th = np.sort(rng.uniform(0, 2 * np.pi, M))
    rad = (50 + 400 * rng.random()) * (1 + 0.3 * np.sin(3 * th + rng.random()) + 0.02 * rng.standard_normal(M))
    pts = np.stack([rad * np.cos(th), 0.6 * rad * np.sin(th)], -1) * scale
    c = np.vstack([pts, pts[:1]])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        (t, (rx, ry), k), u = splprep([c[:, 0], c[:, 1]], s=0.0, k=3, per=1)
    hu, hcx, hcy, st = H.fit_points(pts[None], -108)
    assert not st.any() and np.array_equal(hu[0], u)
    assert np.array_equal(hcx[0], rx) and np.array_equal(hcy[0], ry)
    ot, ocx, ocy = O.fit_periodic_cubic(pts)
    assert np.array_equal(ot, t) and np.array_equal(ocx, rx) and np.array_equal(ocy, ry)


def test_fit_fitpack_small_and_random_lines():
    """FITPACK restatement vs SciPy itself (run here as the checker): bit for bit from the smallest closed line
    (M = 3, where the periodic wrap folds back onto the border block) upwards."""
    import warnings
    from scipy.interpolate import splprep
    rng = np.random.default_rng(3)
    for M in (3, 4, 5, 6, 7, 10, 33, 100, 300):
        for rep in range(6):
            th = np.sort(rng.uniform(0, 2 * np.pi, M))
            pts = np.stack([np.cos(th) * (1 + 0.2 * rng.standard_normal(M)), 0.7 * np.sin(th) + 0.1 * rng.standard_normal(M)], -1)
            c = np.vstack([pts, pts[:1]])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                (t, (rx, ry), k), u = splprep([c[:, 0], c[:, 1]], s=0.0, k=3, per=1)
            hu, hcx, hcy, st = H.fit_points(pts[None], -101)
            assert not st.any() and np.array_equal(hu[0], u)
            assert np.array_equal(hcx[0], rx) and np.array_equal(hcy[0], ry), (M, rep)
            ot, ocx, ocy = O.fit_periodic_cubic(pts)
            assert np.array_equal(ot, t) and np.array_equal(ocx, rx) and np.array_equal(ocy, ry), (M, rep)


@pytest.mark.parametrize("case", ["nan_radius", "nan_position", "all_nan", "nan_mixed"])
def test_qss_nan_inputs_equal_oracle(case):
    """Both schedule kernels (host build) on NaN inputs: same status, same step / iteration counts and the same
    speed / acceleration arrays (NaN pattern included) as the oracle - not just 'status != 0'."""
    d = golden("sim_s10k3_i2")
    ov, hv = O.make_vehicle(*veh_args(d)), H.make_vehicle(*veh_args(d))
    from helpers import nan_cases
    x, y, r, sb = nan_cases()[case]
    o = O.qss(x, y, r, sb, ov, 0)
    assert o["status"] == 2                                   # STO_ORACLE_NAN
    assert o["iters"] < 4 * len(x)                            # terminates like a healthy line, far below the 64 N cap
    for impl in (0, 1, 104, 108):
        h = H.qss(impl, x[None], y[None], r[None], sb, hv)
        assert h["status"][0] == 2, (impl, h["status"][0])    # STO_CAND_NAN, nothing else
        assert h["summary"][0, 6] == o["steps"] and h["summary"][0, 7] == o["iters"] + 1, impl
        for k in ("v", "a"):
            assert np.array_equal(h[k][0], o[k], equal_nan=True), (impl, k)
        assert np.isnan(h["lap"][0]) and np.isnan(o["lap"])


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_out_of_order_list_walk_model_is_exact(lanes):
    """Host model (tests/hostsim/memo3_proto.h) of the out-of-order walk of the re-spawned lists that the lane-group kernel
    sto_qss_memo3.cuh implements: a list sub-pass is cut into chunks of <= 32 open entries; an entry is processed once the
    latest earlier open entry on its own sample and on the sample that writes its source are finished, and it may not
    commit before the earlier entry whose source it writes unless both go in the same round.  Must equal the oracle bit for
    bit (state, segment times, lap, step and iteration counts) on the goldens, random tracks, word-boundary sizes and the
    NaN cases."""
    from helpers import synthetic_closed_track, nan_cases
    H.lib().hostsim_mq_config(32, lanes, 1)
    d = golden("sim_s10k3_i2")
    ov, hv = O.make_vehicle(*veh_args(d)), H.make_vehicle(*veh_args(d))
    cases = []
    for name in ("sim_s10k3_i2", "sim_s30k5_i3", "sim_oval_bank12", "sim_s10k3_i5"):
        g = golden(name)
        cases.append((g["in_X"], g["in_Y"], g["in_CURVATURE"], np.sin(g["in_BANK"])))
    for seed in range(10):
        n = int(np.random.default_rng(900 + seed).integers(128, 1500))
        cases.append(synthetic_closed_track(3000 + seed, n))
    for n in (128, 129, 192, 193, 257):
        idx = np.linspace(0, len(d["in_X"]) - 1, n, endpoint=False).astype(int)
        cases.append((d["in_X"][idx].copy(), d["in_Y"][idx].copy(), d["in_CURVATURE"][idx].copy(), np.zeros(n)))
    for x, y, r, sb in cases:
        o = O.qss(x, y, r, sb, ov, 0)
        h = H.qss(301, x[None], y[None], r[None], sb, hv)
        assert h["status"][0] == 0
        for k in ("v", "a", "lat", "time"):
            assert np.array_equal(h[k][0], o[k]), k
        assert h["lap"][0] == o["lap"] and h["summary"][0, 6] == o["steps"] and h["summary"][0, 7] == o["iters"] + 1
    for case, (x, y, r, sb) in nan_cases().items():
        o = O.qss(x, y, r, sb, ov, 0)
        h = H.qss(301, x[None], y[None], r[None], sb, hv)
        assert h["status"][0] == 2, case
        assert h["summary"][0, 6] == o["steps"] and h["summary"][0, 7] == o["iters"] + 1, case
        for k in ("v", "a"):
            assert np.array_equal(h[k][0], o[k], equal_nan=True), (case, k)


def test_parallel_forward_rows_model_is_exact():
    """Host model (tests/hostsim/memo3_proto.h, memo_forward_rows_par) of the run-parallel forward sub-pass over the original
    rows that the lane-group kernel offers (sto_qss_memo2.cuh, FP): per round the lowest words with dirty fronts each offer
    their lowest dirty front if no dirty front sits below it in its run of consecutive live rows; all are evaluated on the
    state as it stands and committed together.  Must equal the oracle bit for bit; reports rounds against evaluations."""
    import ctypes as C
    from helpers import synthetic_closed_track, nan_cases
    L = H.lib()
    L.hostsim_mq_config(32, 8, 1)
    L.hostsim_mf_config(1)
    out = (C.c_longlong * 3)()
    try:
        d = golden("sim_s10k3_i2")
        ov, hv = O.make_vehicle(*veh_args(d)), H.make_vehicle(*veh_args(d))
        cases = []
        for name in ("sim_s10k3_i2", "sim_s30k5_i3", "sim_oval_bank12", "sim_s10k3_i5"):
            g = golden(name)
            cases.append((g["in_X"], g["in_Y"], g["in_CURVATURE"], np.sin(g["in_BANK"])))
        for seed in range(10):
            n = int(np.random.default_rng(1900 + seed).integers(128, 1500))
            cases.append(synthetic_closed_track(4000 + seed, n))
        for n in (128, 129, 192, 193, 257):
            idx = np.linspace(0, len(d["in_X"]) - 1, n, endpoint=False).astype(int)
            cases.append((d["in_X"][idx].copy(), d["in_Y"][idx].copy(), d["in_CURVATURE"][idx].copy(), np.zeros(n)))
        L.hostsim_mf_stats(out, 1)
        for x, y, r, sb in cases:
            o = O.qss(x, y, r, sb, ov, 0)
            h = H.qss(301, x[None], y[None], r[None], sb, hv)
            assert h["status"][0] == 0
            for k in ("v", "a", "lat", "time"):
                assert np.array_equal(h[k][0], o[k]), k
            assert h["lap"][0] == o["lap"] and h["summary"][0, 6] == o["steps"] and h["summary"][0, 7] == o["iters"] + 1
        L.hostsim_mf_stats(out, 1)
        assert out[1] > 2 * out[0] > 0          # more than two evaluations per round
        for case, (x, y, r, sb) in nan_cases().items():
            o = O.qss(x, y, r, sb, ov, 0)
            h = H.qss(301, x[None], y[None], r[None], sb, hv)
            assert h["status"][0] == 2, case
            assert h["summary"][0, 6] == o["steps"] and h["summary"][0, 7] == o["iters"] + 1, case
            for k in ("v", "a"):
                assert np.array_equal(h[k][0], o[k], equal_nan=True), (case, k)
    finally:
        L.hostsim_mf_config(0)


def test_list_by_runs_of_rows_model_is_exact():
    """Host model (tests/hostsim/memo3_proto.h, memo_spawned_rows_runs) of the next step for the re-spawned lists: the live
    entries occupy virtual rows, an entry only affects entries on the neighbouring row, so a maximal run of occupied
    adjacent rows - cut at its first row whose edge memo is not CONT - is independent of every other run and its entries
    are processed in list order.  Must equal the oracle bit for bit; on the Monza line the forward list then touches
    ~11 k entries instead of ~44 k and needs ~1.3 k rounds of 8 lanes for its 5.4 k evaluations."""
    import ctypes as C
    from helpers import synthetic_closed_track
    L = H.lib()
    L.hostsim_mq_config(32, 8, 1)
    L.hostsim_mr_config(1)
    out = (C.c_longlong * 14)()
    try:
        d = golden("sim_s10k3_i2")
        ov, hv = O.make_vehicle(*veh_args(d)), H.make_vehicle(*veh_args(d))
        cases = []
        for name in ("sim_s10k3_i2", "sim_s30k5_i3", "sim_oval_bank12"):
            g = golden(name)
            cases.append((g["in_X"], g["in_Y"], g["in_CURVATURE"], np.sin(g["in_BANK"])))
        for seed in range(6):
            n = int(np.random.default_rng(2900 + seed).integers(128, 1500))
            cases.append(synthetic_closed_track(5000 + seed, n))
        L.hostsim_mr_stats(out, 1)
        for x, y, r, sb in cases:
            o = O.qss(x, y, r, sb, ov, 0)
            h = H.qss(301, x[None], y[None], r[None], sb, hv)
            assert h["status"][0] == 0
            for k in ("v", "a", "lat", "time"):
                assert np.array_equal(h[k][0], o[k]), k
            assert h["lap"][0] == o["lap"] and h["summary"][0, 6] == o["steps"] and h["summary"][0, 7] == o["iters"] + 1
        L.hostsim_mr_stats(out, 1)
        fwd = dict(zip(("walks", "runs", "evals", "rounds_est", "longest", "entries_in_runs", "visits"), out[7:]))
        assert 0 < fwd["entries_in_runs"] < fwd["visits"] // 2 and fwd["rounds_est"] < fwd["evals"] // 2
    finally:
        L.hostsim_mr_config(0)
