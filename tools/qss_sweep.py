"""Developer tool (GPU): time the fused lap-time path for several batch sizes / candidates-per-warp settings."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spline_trajectory_optimization_b200 import _lib  # noqa: E402
from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator  # noqa: E402

lib = _lib.load()
rt, veh = bench.build_track(), bench.test_vehicle()
configs = [tuple(int(x) for x in a.split(":")[:2]) + (a.split(":")[2] if a.count(":") > 1 else "s", a.split(":")[3] if a.count(":") > 2 else "4", a.split(":")[4] if a.count(":") > 3 else "0") for a in sys.argv[1:]] or [(4096, 8, 's', '4', '0')]
cache = {}
ref = None
for B, lanes, planes, group, kern in configs:
    if B not in cache:
        ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh, impl="memo")
        off = bench.make_offsets(rt, min(B, 4096), 1234)
        if B > 4096:
            off = np.tile(off, (B // 4096, 1))
        cache[B] = (ev, ev.to_sample_major(torch.from_numpy(off).cuda()))
    ev, d_off = cache[B]
    _lib.check(lib.sto_set_tuning(b"qss_lanes", int(lanes)))
    _lib.check(lib.sto_set_tuning(b"qss_planes", {"s": 1, "g": 2, "g0": 3, "g1": 4}[planes]))
    _lib.check(lib.sto_set_tuning(b"qss_group", int(group)))
    _lib.check(lib.sto_set_tuning(b"qss_kernel", int(kern)))
    lib.sto_set_stage_timing(1)
    best = None
    for it in range(3):
        lap, st = ev.lap_times(d_off, B=B)
        buf = (ctypes.c_float * 4)()
        _lib.check(lib.sto_last_stage_ms(buf))
        ms = list(buf)
        if best is None or ms[3] < best[3]:
            best = ms
    lap = lap.cpu().numpy()
    if ref is None:
        ref = lap[:4096].copy()
    same = np.array_equal(lap[:min(B, 4096)], ref[:min(B, 4096)])
    print(f"B={B:7d} lanes={lanes:2d} planes={planes} group={group} kernel={kern}  fit {best[1]:8.2f} ms  sample {best[2]:8.2f} ms  qss {best[3]:9.2f} ms  "
          f"-> {B / (sum(best) * 1e-3):10.0f} cand/s   laps identical to first config: {same}  status ok: {not st.any().item()}",
          flush=True)
