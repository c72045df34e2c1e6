"""Developer tool (GPU): fit kernel time on a long track (Monza at 0.25 m, M = 23,160) - one-lane Thomas solve vs the
partitioned lane-group solve - for several batch sizes.  python tools/fit_part_bench.py [interval_m]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import test_vehicle_params, to_sm  # noqa: E402
from spline_trajectory_optimization_b200 import _lib, candidates, tracks  # noqa: E402
from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator  # noqa: E402
from spline_trajectory_optimization_b200.models.race_track import RaceTrack  # noqa: E402
from spline_trajectory_optimization_b200.models.vehicle import Vehicle  # noqa: E402

interval = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
lib = _lib.load()
c, l, r = tracks.monza_raw()
rt = RaceTrack("monza", l, r, c, s=10.0, interval=interval)
M = len(rt.center_d)
ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), Vehicle(test_vehicle_params()))
base = candidates.smooth_offsets(M, 64, rt.dist_to_left, rt.dist_to_right, seed=3)
if os.environ.get("SWEEP"):   # explicit lane counts: python tools/fit_part_bench.py 0.25 with SWEEP=1
    for B in [int(x) for x in os.environ.get("SWEEP_B", "256,1024,4096,16384").split(",")]:
        off = to_sm(np.tile(base, ((B + 63) // 64, 1))[:B])
        line = f"M={M} B={B:6d}"
        for mode, lanes in ((2, 1), (2, 4), (2, 8), (2, 16), (2, 32), (0, 1), (0, 2), (0, 8), (1, 1), (1, 2), (1, 4), (1, 8), (1, 16), (1, 32)):
            lib.sto_set_tuning(b"fit_split", lanes)
            lib.sto_set_fit_solver(mode)
            best = 1e9
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ev.fit(off, B=B)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            line += f"  {('thomas', 'blocks', 'fitpack')[mode]}x{lanes}: {best:8.2f} ms"
        print(line, flush=True)
    lib.sto_set_tuning(b"fit_split", 0)
    lib.sto_set_fit_solver(2)
    sys.exit(0)
for B in (1, 64, 1024, 4096, 16384):
    off = to_sm(np.tile(base, ((B + 63) // 64, 1))[:B])
    res = {}
    for mode, name in ((0, "one-lane"), (1, "partitioned"), (2, "fitpack")):
        lib.sto_set_fit_solver(mode)
        lanes = lib.sto_fit_solver_lanes(M, B)
        best = 1e9
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            u, cx, cy, st = ev.fit(off, B=B)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[name] = (best, lanes, cx[:, :B].clone())
    lib.sto_set_fit_solver(2)
    a, b, f = res["one-lane"], res["partitioned"], res["fitpack"]
    err = float(((f[2] - b[2]).abs().max() / f[2].abs().max()).item())
    print(f"M={M} B={B:6d}  fitpack {f[0]:9.3f} ms   thomas {a[0]:9.3f} ms   blocks ({b[1]:2d} lanes) {b[0]:9.3f} ms   "
          f"max |dc|/|c| blocks vs fitpack {err:.2e}", flush=True)
