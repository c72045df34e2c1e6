"""FAST MODE (csrc/sto_fast.cuh, sto_lap_time_fast_f64): the separately-reported streaming variant of the path - Thomas
fit + sampler + two-sweep QSS with the reference's step operator + fill_time in one kernel.

It is NOT the reference's algorithm (Simulator.run_simulation's lock-step multi-front schedule is not emulated), so two
different things are tested:
  * the kernel against ITS OWN CPU twin (oracle: Thomas fit -> _deBoor_D sampler -> sto_oracle_qss_sweep), bit for bit
    (CPU: the host build of the device functions; GPU: through the C ABI);
  * its DEVIATION from the reference, measured on the goldens of the unmodified reference and bounded (not a parity
    claim): |lap difference| <= 3.5 s, reported per case.
"""
import numpy as np
import pytest

import oracle_py as O
from helpers import CAND_CASES, SIM_CASES, golden, veh_args


def _twin_laps(d, rounds, sinb=None):
    """The fast path restated with the oracle's pieces: Thomas fit, sampler, sweep QSS (ref_pow = 0: x*x arithmetic)."""
    ov = O.make_vehicle(*veh_args(d))
    sb = np.zeros(len(d["ts"])) if sinb is None else sinb
    O.set_fit_solver("thomas")
    try:
        laps, outs = [], []
        for pts in d["points"]:
            t, cx, cy = O.fit_periodic_cubic(pts)
            X, Y, YAW, R = O.sample(t, cx, cy, 3, d["ts"], 0)
            o = O.qss_sweep(X, Y, R, sb, ov, rounds, 0)
            laps.append(o["lap"])
            outs.append(o)
    finally:
        O.set_fit_solver("fitpack")
    return np.array(laps), outs


@pytest.mark.parametrize("name", CAND_CASES[:2])
@pytest.mark.parametrize("rounds", [1, 2])
def test_fast_path_host_build_equals_its_cpu_twin(name, rounds):
    import hostsim_py as H
    d = golden(name)
    hv = H.make_vehicle(*veh_args(d))
    nx, ny = -np.sin(d["centre_yaw"]), np.cos(d["centre_yaw"])
    r = H.fast(d["centre_x"], d["centre_y"], nx, ny, None, d["ts"], d["offsets"], hv, rounds)
    laps, outs = _twin_laps(d, rounds)
    assert not r["status"].any()
    assert np.array_equal(r["lap"], laps)
    for b, o in enumerate(outs):
        assert np.array_equal(r["v"][b], o["v"]) and np.array_equal(r["a"][b], o["a"])
        assert np.array_equal(r["time"][b], o["time"])


@pytest.mark.parametrize("name", SIM_CASES + ["sim_s30k5_i1"])
def test_sweep_deviation_from_the_reference_is_bounded(name):
    """How far the two-sweep result is from the reference's schedule (goldens of the unmodified reference): laps within
    3.5 s (measured 0.005 .. 3.0 s; the sweep is the faster lap - the reference's schedule artefacts cost time), and the
    sweep has practically converged after 3 rounds (rounds 3 and 4 within a millisecond)."""
    d = golden(name)
    ov = O.make_vehicle(*veh_args(d))
    sb = np.sin(d["in_BANK"])
    laps = [O.qss_sweep(d["in_X"], d["in_Y"], d["in_CURVATURE"], sb, ov, r, 1)["lap"] for r in (1, 2, 3, 4)]
    ref = float(d["lap"])
    assert abs(laps[1] - ref) < 3.5, (name, laps[1] - ref)
    assert laps[1] < ref + 0.05                      # never meaningfully slower than the reference's profile
    assert abs(laps[3] - laps[2]) < 1e-3
    assert abs(laps[1] - laps[3]) < 0.1              # two rounds are within 0.1 s of the sweep's fixed point


@pytest.mark.gpu
@pytest.mark.parametrize("name", CAND_CASES)
@pytest.mark.parametrize("stage", [0, -1])
def test_fast_kernel_equals_its_cpu_twin_on_device(name, stage):
    """GPU, through the C ABI: laps and full outputs of sto_lap_time_fast_f64 == the CPU twin, bit for bit, with the track
    tables staged in shared memory by the bulk-copy engine (stage = -1) and read from global memory (stage = 0)."""
    import torch
    from helpers import to_cm, to_sm
    from spline_trajectory_optimization_b200 import _lib
    from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator
    d = golden(name)
    nrm = np.stack([-np.sin(d["centre_yaw"]), np.cos(d["centre_yaw"])], axis=1)
    ctr = np.stack([d["centre_x"], d["centre_y"]], axis=1)
    ev = BatchedLineEvaluator(ctr, nrm, d["ts"], _lib.make_vehicle(*veh_args(d)))
    B = d["offsets"].shape[0]
    for rounds in (1, 2):
        lap, st, out = ev.lap_times_fast(to_sm(d["offsets"]), B=B, rounds=rounds, outputs=True, stage_tables=stage)
        torch.cuda.synchronize()
        laps, outs = _twin_laps(d, rounds)
        assert not st.cpu().numpy().any()
        assert np.array_equal(lap.cpu().numpy(), laps), (rounds, lap.cpu().numpy() - laps)
        for b, o in enumerate(outs):
            assert np.array_equal(to_cm(out["speed"], B)[b], o["v"]) and np.array_equal(to_cm(out["time"], B)[b], o["time"])
            assert np.array_equal(to_cm(out["lat_acc"], B)[b], o["lat"])
        lap2, st2 = ev.lap_times_fast(to_sm(d["offsets"]), B=B, rounds=rounds)       # lap-only: same laps
        assert torch.equal(lap2, lap)
    # deviation from the exact schedule on the same lines: bounded, and the ranking signal is there
    exact, _ = ev.lap_times(to_sm(d["offsets"]), B=B)
    fast, _ = ev.lap_times_fast(to_sm(d["offsets"]), B=B, rounds=2)
    dev = (fast - exact).cpu().numpy()
    assert np.max(np.abs(dev)) < 3.5, dev


@pytest.mark.gpu
def test_fast_kernel_banked_and_ragged(sto=None):
    """Banked oval (sin(bank) table staged too), a ragged batch of 37 and B = 1."""
    import torch
    from helpers import test_vehicle_params, to_sm
    from spline_trajectory_optimization_b200 import candidates
    from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator
    from spline_trajectory_optimization_b200.models.race_track import RaceTrack
    from spline_trajectory_optimization_b200.models.trajectory import Trajectory
    from spline_trajectory_optimization_b200.models.vehicle import Vehicle
    centre, left, right = candidates.banked_oval(straight=300.0, radius=120.0, spacing=10.0)
    rt = RaceTrack("oval", left, right, centre, s=1.0, interval=4.0)
    M = len(rt.center_d)
    bank = rt.center_d[:, Trajectory.BANK]
    off = candidates.smooth_offsets(M, 37, rt.dist_to_left, rt.dist_to_right, seed=5)
    nrm = rt.left_normals()
    ev = BatchedLineEvaluator(rt.center_d[:, :2], nrm, rt.center_d.ts(), Vehicle(test_vehicle_params()), bank=bank)
    d = dict(points=np.stack([rt.center_d[:, :2] + off[b][:, None] * nrm for b in range(37)]), ts=rt.center_d.ts())
    g = golden("cand_m579_n579")
    d.update({k: g[k] for k in ("veh_scalars", "veh_acc_x", "veh_acc_c", "veh_dcc_x", "veh_dcc_c")})
    laps, _ = _twin_laps(d, 2, sinb=np.sin(bank))
    for stage in (0, -1):
        lap, st = ev.lap_times_fast(to_sm(off), B=37, rounds=2, stage_tables=stage)
        assert not st.cpu().numpy().any() and np.array_equal(lap.cpu().numpy(), laps)
        lap1, st1 = ev.lap_times_fast(to_sm(off[5:6]), B=1, rounds=2, stage_tables=stage)
        assert float(lap1[0]) == laps[5]
