"""Developer tool (GPU): per-phase SM-clock breakdown of qss_memo_kernel.  Builds a second copy of the library with
-DSTO_PHASE_CLOCKS (clock64() deltas accumulated per thread, returned through the summary rows) and runs the
BASELINE config-2 batch through the fit/sample/QSS stages.  Not part of the product."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spline_trajectory_optimization_b200 import _lib, build  # noqa: E402

# built here (the authoring container cross-compiles; `--build-only`) and shipped to the GPU box with the tree
alt = os.path.join(ROOT, "ab", "libsto_b200_phase.so")
os.makedirs(os.path.dirname(alt), exist_ok=True)
if "--build-only" in sys.argv or not os.path.exists(alt):
    cmd = [build.nvcc_path()] + [f for f in build.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DSTO_PHASE_CLOCKS", "-o", alt,
           os.path.join(build.CSRC, "sto_b200.cu")]
    subprocess.check_call(cmd)
    if "--build-only" in sys.argv:
        sys.exit(0)
_lib.LIB_PATH = alt
import bench  # noqa: E402
from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator, run_qss  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
rt, veh = bench.build_track(), bench.test_vehicle()
ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh, impl="memo")
off = ev.to_sample_major(torch.from_numpy(bench.make_offsets(rt, B, 1234)).cuda())
u, cx, cy, st = ev.fit(off, B=B)
S = ev.sample(u, cx, cy, B=B, want=("x", "y", "radius"))
for _ in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = run_qss(S["x"], S["y"], S["radius"], ev.vehicle, B=B, impl=_lib.QSS_MEMO, profiles=True)
    e1.record()
    torch.cuda.synchronize()
    print("qss ms (incl. chord kernel)", e0.elapsed_time(e1))
summ = res["summary"][:, :B].cpu().numpy()
sub = res["lat_acc"][:8, :B].cpu().numpy()
if int(os.environ.get("STO_QSS_KERNEL", "0")) == 1:
    names = ["loop head", "orig backward", "spawned backward", "orig forward", "spawned forward", "compaction",
             "loop tail", "finish"]
    tot = summ.sum(axis=0)
    print("per-thread total clocks: mean %.3g  (%.1f ms at 1.965 GHz)" % (tot.mean(), tot.mean() / 1.965e6))
    for k, n in enumerate(names):
        print("%-18s %6.2f %%   mean clocks %.3g" % (n, 100 * summ[k].mean() / tot.mean(), summ[k].mean()))
    subn = ["orig search clk", "orig eval clk", "spawned search clk", "spawned eval clk", "orig evals (lane)", "spawned evals (lane)",
            "spawned warp rounds", "orig warp rounds"]
    for k, n in enumerate(subn):
        print("%-22s mean %.4g   max %.4g" % (n, sub[k].mean(), sub[k].max()))
else:
    # one-loop kernel (sto_qss_memo2.cuh): [4 k + phase], k = 0 search, 1 evaluate, 2 commit + post clocks, 3 warp rounds
    allv = np.concatenate([summ, sub], axis=0)
    ph = ["bwd original rows", "bwd re-spawned list", "fwd original rows", "fwd re-spawned list"]
    tot = allv[:12].sum(axis=0).mean()
    print("clocks inside the round loop per candidate: mean %.3g (%.1f ms at 1.965 GHz)" % (tot, tot / 1.965e6))
    for p_ in range(4):
        se, ev, co, rd = (allv[4 * k + p_].mean() for k in range(4))
        print("%-20s search %5.1f %%  evaluate %5.1f %%  commit+post %5.1f %%   rounds %7.0f   clocks/round: search %6.0f eval %6.0f commit %6.0f"
              % (ph[p_], 100 * se / tot, 100 * ev / tot, 100 * co / tot, rd, se / max(rd, 1), ev / max(rd, 1), co / max(rd, 1)))
    oth = res["time"][:8, :B].cpu().numpy()
    names3 = ["set-up (records, planes)", "phase entry (word scans, sweep0, list heads)", "fold + iteration tail", "finish (fill_time, lap)",
              "whole kernel", "record loads in flight (all evaluations of this lane)", "evaluations redone with the plain operators (count)",
              "arithmetic of the evaluations (clocks, this lane)"]
    for k, n in enumerate(names3):
        print("%-46s mean %.3g  max %.3g clocks  (%.2f / %.2f ms)" % (n, oth[k].mean(), oth[k].max(), oth[k].mean() / 1.965e6, oth[k].max() / 1.965e6))
    print("round loop, max over candidates: %.2f ms" % (allv[:12].sum(axis=0).max() / 1.965e6))
    if int(os.environ.get("STO_QSS_KERNEL", "0")) == 3:
        c4 = res["lon_acc"][:8, :B].cpu().numpy().mean(axis=1)
        print("list phases (sto_qss_memo3.cuh): scan %.3g clocks in %.0f steps (%.0f per step), selection %.3g clocks in %.0f passes "
              "(%.0f per pass), chunk end %.3g clocks for %.0f chunks (%.0f per chunk)" % (
                  c4[0], c4[1], c4[0] / max(c4[1], 1), c4[2], c4[3], c4[2] / max(c4[3], 1), c4[4], c4[5], c4[4] / max(c4[5], 1)))
