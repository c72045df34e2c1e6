"""Developer tool (GPU): time the fast mode for several builds of the library.  usage: fast_ab.py lib1.so [lib2.so ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if sys.argv[1] != "--child":
    for lib in sys.argv[1:]:
        out = subprocess.run([sys.executable, __file__, "--child", lib], capture_output=True, text=True)
        print(lib, (out.stdout.strip().splitlines() or [out.stderr[-400:]])[-1], flush=True)
    sys.exit(0)
sys.path.insert(0, ROOT)
from spline_trajectory_optimization_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[2])
import torch, bench
from spline_trajectory_optimization_b200 import candidates
from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator
rt, veh = bench.build_track(), bench.test_vehicle()
ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh)
res = []
for B in (4096, 131072):
    d = candidates.smooth_offsets_device(ev.M, 0, B, rt.dist_to_left, rt.dist_to_right, ev.device, seed=77)
    for kw in (dict(stage_tables=0), dict(stage_tables=0, outputs=True)):
        r = ev.lap_times_fast(d, B=B, rounds=2, **kw)
        if kw.get("outputs"):
            kw["outputs"] = r[2]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = ev.lap_times_fast(d, B=B, rounds=2, **kw)
        e1.record()
        torch.cuda.synchronize()
        res.append("%s%s %.2f ms" % (B, "/out" if "outputs" in kw else ("/staged" if kw["stage_tables"] else "/global"), e0.elapsed_time(e1) / 3))
    lap0 = float(r[0][0])
    del d, r
print("  ".join(res), " lap0 %.9f" % lap0)
