"""Developer tool (GPU): A/B two builds of the library on the fused path.  usage: ab_lib.py libA.so libB.so [B]"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 3 and sys.argv[1] != "--child":
    for lib in sys.argv[1:3]:
        out = subprocess.run([sys.executable, __file__, "--child", lib, sys.argv[3]], capture_output=True, text=True)
        print(lib, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
    sys.exit(0)
sys.path.insert(0, ROOT)
from spline_trajectory_optimization_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[2])
import numpy as np, torch, bench
from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator
B = int(sys.argv[3])
lib = _lib.load()
rt, veh = bench.build_track(), bench.test_vehicle()
ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh)
d_off = ev.to_sample_major(torch.from_numpy(bench.make_offsets(rt, B, 1234)).cuda())
lib.sto_set_stage_timing(1)
best = None
for it in range(4):
    lap, st = ev.lap_times(d_off, B=B)
    buf = (ctypes.c_float * 4)(); _lib.check(lib.sto_last_stage_ms(buf)); ms = list(buf)
    if best is None or ms[3] < best[3]: best = ms
print(f"fit {best[1]:.2f} sample {best[2]:.2f} qss {best[3]:.2f} ms  lap0 {float(lap[0]):.9f} ok {not st.any().item()}")
