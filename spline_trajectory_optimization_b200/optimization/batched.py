"""Evaluate many control-point variants of one racing-line spline in a single launch.

The reference's ``TrajectoryOptimizer`` moves one control point (or a 5-point window) at a time and, after every QP
step, re-samples the spline and re-runs the QSS to score it (``optimization/optimizer.py:196-211`` joint window,
``:276-289,333-338`` per point).  Its QP solves (qpOASES through CasADi) stay on the host and are out of scope; what
this module batches is the scoring: all candidate edits of a sweep are stacked as coefficient sets on the spline's
shared knot vector and go through ``sto_lap_time_splines_f64`` together.
"""
import numpy as np


def wrap_control_points(c, k):
    """Periodic closure of a degree-k closed spline's coefficient vector, as the reference re-imposes it after every
    edit (optimizer.py:201-205 for k = 5): the first k//2 coefficients are copied from the tail block, the last
    k - k//2 from the head block, i.e. c[n-k+i] == c[i]."""
    n = len(c)
    front = k // 2
    for i in range(front):
        c[i] = c[n - k + i]
    for i in range(front, k):
        c[n - k + i] = c[i]
    return c


def control_point_variants(spline, edits):
    """spline: BSplineTrajectory; edits: list of dicts {control_point_index: (x, y)}.  Returns (cx[B, n], cy[B, n]):
    candidate 0.. as the reference would hold them after ``set_control_point`` + the periodic wrap copies."""
    k = int(spline._spl_x.k)
    base_x, base_y = np.asarray(spline._spl_x.c, dtype=np.float64), np.asarray(spline._spl_y.c, dtype=np.float64)
    cx = np.repeat(base_x[None, :], len(edits), axis=0)
    cy = np.repeat(base_y[None, :], len(edits), axis=0)
    for b, edit in enumerate(edits):
        for idx, (x, y) in edit.items():
            cx[b, idx] = x
            cy[b, idx] = y
        wrap_control_points(cx[b], k)
        wrap_control_points(cy[b], k)
    return cx, cy


def score_control_point_variants(spline, edits, ts, vehicle, bank=None, device=None):
    """Lap time (sum of TIME) of every variant: spline.sample_along(ts=ts) -> run_simulation, batched on the GPU.
    Returns (lap[B], status[B]) as NumPy arrays."""
    import torch
    from .. import evaluator
    if not torch.cuda.is_available():
        raise RuntimeError("score_control_point_variants needs a CUDA device (there is no CPU fallback)")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    cx, cy = control_point_variants(spline, edits)
    B = cx.shape[0]
    ld = evaluator.round_up32(B)

    def sm(a):
        out = torch.zeros((a.shape[1], ld), dtype=torch.float64, device=dev)
        out[:, :B] = torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
        if ld > B:   # padding lanes carry a valid spline (they are evaluated, their results dropped)
            out[:, B:] = out[:, :1]
        return out

    sinb = None if bank is None else np.sin(np.asarray(bank, dtype=np.float64))
    lap, st = evaluator.lap_times_splines(spline._spl_x.t, spline._spl_x.k, sm(cx), sm(cy), ts, vehicle, B=B,
                                          sin_bank=sinb)
    return lap.cpu().numpy(), st.cpu().numpy()
