"""Developer tool (GPU): where the host-buffer entry point spends its time."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator
rt, veh = bench.build_track(), bench.test_vehicle()
B = 4096
ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh)
off = bench.make_offsets(rt, B, 1234)
pin = torch.from_numpy(off).pin_memory()
dev = torch.empty_like(pin, device="cuda")
for _ in range(3):
    dev.copy_(pin, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    dev.copy_(pin, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"pinned H2D of {pin.numel()*8/1e6:.1f} MB: {dt*1e3:.2f} ms  ({pin.numel()*8/dt/1e9:.1f} GB/s)")
lap = torch.empty(B, dtype=torch.float64).pin_memory(); st = torch.empty(B, dtype=torch.int32).pin_memory()
for _ in range(3):
    ev.lap_times_host_into(pin.data_ptr(), B, lap.data_ptr(), st.data_ptr())
t0 = time.perf_counter()
for _ in range(10):
    ev.lap_times_host_into(pin.data_ptr(), B, lap.data_ptr(), st.data_ptr())
dt = (time.perf_counter() - t0) / 10
print(f"sto_lap_time_host_f64: {dt*1e3:.2f} ms per call")
d_off = ev.to_sample_major(dev)
for _ in range(3):
    ev.lap_times(d_off, B=B)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    ev.lap_times(d_off, B=B)
torch.cuda.synchronize()
print(f"sto_lap_time_f64 (device resident): {(time.perf_counter()-t0)/10*1e3:.2f} ms per call")
