"""Developer tool (GPU): the fast mode at one batch size (for ncu captures).  usage: fast_probe.py [B] [outputs 0/1]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spline_trajectory_optimization_b200 import candidates
from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
outputs = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
rt, veh = bench.build_track(), bench.test_vehicle()
ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh)
d = candidates.smooth_offsets_device(ev.M, 0, B, rt.dist_to_left, rt.dist_to_right, ev.device, seed=77)
for _ in range(3):
    r = ev.lap_times_fast(d, B=B, rounds=2, outputs=outputs)
torch.cuda.synchronize()
print("ok", float(r[0][0]), bool(r[1].any()))
