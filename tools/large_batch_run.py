"""Developer tool (GPU): one fused pass at a large batch (for ncu captures of the global-plane kernel)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rt, veh = bench.build_track(), bench.test_vehicle()
ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh)
off = bench.make_offsets(rt, 4096, 1234)
d_off = ev.to_sample_major(torch.from_numpy(np.tile(off, (B // 4096, 1))).cuda())
for _ in range(2):
    lap, st = ev.lap_times(d_off, B=B)
torch.cuda.synchronize()
print("ok", float(lap[0]), bool(st.any()))
