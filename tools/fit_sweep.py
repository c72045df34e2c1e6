"""Developer tool (GPU): fit kernel time vs lanes per candidate."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spline_trajectory_optimization_b200 import _lib
from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator
lib = _lib.load()
rt, veh = bench.build_track(), bench.test_vehicle()
for B in (4096, 32768):
    ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh)
    off = bench.make_offsets(rt, 4096, 1234)
    d_off = ev.to_sample_major(torch.from_numpy(np.tile(off, (B // 4096, 1))).cuda())
    ref = None
    for split in (1, 2, 4, 8, 16, 32):
        lib.sto_set_tuning(b"fit_split", split)
        lib.sto_set_stage_timing(1)
        best = 1e9
        for _ in range(3):
            lap, st = ev.lap_times(d_off, B=B)
            buf = (ctypes.c_float * 4)(); _lib.check(lib.sto_last_stage_ms(buf)); best = min(best, buf[1])
        l = lap.cpu().numpy()
        ref = l if ref is None else ref
        print(f"B={B} fit split={split:2d}: {best:.2f} ms  laps identical: {np.array_equal(l, ref)}", flush=True)
