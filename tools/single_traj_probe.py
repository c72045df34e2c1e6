"""Developer tool (GPU): latency of the single-trajectory drop-in (Simulator.run_simulation) vs trajectory length."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import golden, test_vehicle_params
from spline_traj_optm.models.trajectory import Trajectory
from spline_traj_optm.models.vehicle import Vehicle
from spline_traj_optm.simulator.simulator import Simulator
from spline_trajectory_optimization_b200 import _lib
from spline_trajectory_optimization_b200.evaluator import run_qss
for name in ("sim_s30k5_i10", "sim_s10k3_i2", "sim_s30k5_i1"):
    d = golden(name)
    n = len(d["in_X"])
    traj = Trajectory(n)
    traj[:, 0], traj[:, 1], traj[:, 5], traj[:, 13] = d["in_X"], d["in_Y"], d["in_CURVATURE"], d["in_BANK"]
    sim = Simulator(Vehicle(test_vehicle_params()))
    sim.run_simulation(traj)
    t0 = time.perf_counter(); res = sim.run_simulation(traj); dt = time.perf_counter() - t0
    print(f"{name}: N={n} run_simulation (default: memoised kernel) {dt*1e3:.1f} ms  lap {res.lap_time:.9f} (reference took {float(d['ref_run_time']):.1f} s)")
    dev = torch.device("cuda")
    col = lambda a: torch.zeros((n, 32), dtype=torch.float64, device=dev).index_put_((torch.arange(n, device=dev), torch.zeros(n, dtype=torch.long, device=dev)), torch.from_numpy(np.ascontiguousarray(a)).to(dev))
    x, y, r = col(d["in_X"]), col(d["in_Y"]), col(d["in_CURVATURE"])
    for impl, nm in ((_lib.QSS_MEMO, "memo"), (_lib.QSS_PLAIN, "plain")):
        run_qss(x, y, r, sim.vehicle, B=1, sin_bank=np.sin(d["in_BANK"]), impl=impl); torch.cuda.synchronize()
        t0 = time.perf_counter(); out = run_qss(x, y, r, sim.vehicle, B=1, sin_bank=np.sin(d["in_BANK"]), impl=impl); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"    run_qss B=1 {nm}: {dt*1e3:.1f} ms  lap {float(out['lap'][0]):.9f}")
