#!/usr/bin/env python
"""bench.py -- candidate racing lines / second (fit + curvature + QSS lap time) on Monza.

    python bench.py [--gpus N] [--steps K] [--warmup W]                      # our arm (CUDA, libsto_b200.so)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]     # CPU arm: the oracle port, all host threads
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...         # one rank per GPU, weak scaling

A step = one pass of the hot path over one batch of synthetic candidates:
    offsets[M, B] (resident in HBM)  ->  periodic cubic fit  ->  resample + turn radius  ->  exact-schedule QSS
    -> lap[B]  (+ for N > 1: local argmin, NCCL all-gather of the (lap, index) pairs, global argmin).
Workload at N = 1: BASELINE.json configs[1] -- Monza, 4,096 lateral-offset candidates, 2 m spacing
(M = N = 2895), FP64.  For N > 1 every rank evaluates its own 4,096 candidates (weak scaling).
Prints ONE JSON line on rank 0 (see the task contract for the keys).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CANDIDATES_PER_GPU = 4096
INTERVAL_M = 2.0
METRIC = "candidate racing lines/sec (fit+curvature+QSS lap time), Monza"
UNIT = "candidates/s"


# ----------------------------------------------------------------------------------------------------------------
def build_track(interval=INTERVAL_M):
    from spline_trajectory_optimization_b200 import tracks
    from spline_trajectory_optimization_b200.models.race_track import RaceTrack
    c, l, r = tracks.monza_raw()
    return RaceTrack("monza", l, r, c, s=10.0, interval=interval)


def test_vehicle():
    """Vehicle of the reference's tests/test_simulator.py:16-20."""
    from spline_trajectory_optimization_b200.models.vehicle import Vehicle, VehicleParams
    acc = np.array([[0.0, 10.0], [50.0, 7.0], [100.0, 0.5]])
    dcc = np.array([[0.0, -13.0], [50.0, -15.0], [100.0, -20.0]])
    return Vehicle(VehicleParams(acc, dcc, 10.0, -20.0, 15.0, -15.0, 100.0, 30.0))


def make_offsets(rt, B, seed):
    from spline_trajectory_optimization_b200 import candidates
    return candidates.smooth_offsets(len(rt.center_d), B, rt.dist_to_left, rt.dist_to_right, seed=seed)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([p.strip() for p in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
def oracle_setup(rt, veh):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    O.build()
    p = veh.param
    ov = O.make_vehicle((p.max_lon_acc_mpss, p.max_lon_dcc_mpss, p.max_left_acc_mpss, p.max_right_acc_mpss,
                         p.max_speed_mps, p.max_jerk), veh.acc_intp.x, veh.acc_intp.c, veh.dcc_intp.x, veh.dcc_intp.c)
    return O, ov


def cpu_baseline_leg(rt, veh, offsets, seconds_target=12.0):
    """The oracle port (plain C restatement of the reference's schedule, ref_pow=1 arithmetic) on all host threads
    over a bounded sample of the same candidates.  A reported baseline, not a target."""
    O, ov = oracle_setup(rt, veh)
    cores = os.cpu_count() or 1
    nrm = rt.left_normals()
    c = rt.center_d
    ts, sb = c.ts(), np.zeros(len(c))
    # calibrate on one candidate per thread, then size the sample for ~seconds_target of wall time
    n0 = min(len(offsets), cores)
    t0 = time.perf_counter()
    O.lap_batch(c[:, 0], c[:, 1], nrm[:, 0], nrm[:, 1], offsets[:n0], ts, sb, ov, n_threads=cores, ref_pow=1)
    per_round = max(time.perf_counter() - t0, 1e-3)
    n = int(min(len(offsets), max(n0, cores * max(1, int(seconds_target / per_round)))))
    t0 = time.perf_counter()
    lap, st = O.lap_batch(c[:, 0], c[:, 1], nrm[:, 0], nrm[:, 1], offsets[:n], ts, sb, ov, n_threads=cores, ref_pow=1)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} of the step's candidates, oracle/sto_oracle.c lap_batch on {cores} threads, {dt:.2f} s"}, lap[:n]


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Python
    (~15 s / candidate); the arm times its compiled restatement (oracle/, bit-exact against it) on all host
    threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rt, veh = build_track(), test_vehicle()
    O, ov = oracle_setup(rt, veh)
    cores = os.cpu_count() or 1
    c, nrm = rt.center_d, rt.left_normals()
    ts, sb = c.ts(), np.zeros(len(c))
    per_step = max(cores * 8, 64)
    off = make_offsets(rt, per_step, seed=1234)
    run = lambda: O.lap_batch(c[:, 0], c[:, 1], nrm[:, 0], nrm[:, 1], off, ts, sb, ov, n_threads=cores, ref_pow=1)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lap, st = run()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "Monza 2 m (M=N=2895), lateral-offset candidates, FP64",
                       "candidates_per_step": per_step, "note": "bounded sample of the 4096-candidate batch"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per_step} candidates/step on {cores} threads (oracle/sto_oracle.c)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "lap_min_s": float(np.min(lap))}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from spline_trajectory_optimization_b200 import _lib
    from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU arm has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's banner ("NCCL version ...", printed on stdout when the
        # image exports NCCL_DEBUG=VERSION/INFO) out of it unless explicitly asked for
        if "STO_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["STO_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)   # even WARN prints the version banner
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B = args.candidates
    rt, veh = build_track(), test_vehicle()
    M = N = len(rt.center_d)
    ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh, device=dev, impl=args.qss)
    # two different candidate batches per rank, alternated between steps: no step re-reads its predecessor's data;
    # one step's working set (offsets 95 MB + 1.5 GB of state/workspace) is >> the 126 MB L2 anyway
    host_off = [make_offsets(rt, B, seed=1234 + 7919 * rank + 104729 * j) for j in range(2)]
    d_off = [ev.to_sample_major(torch.from_numpy(o).to(dev)) for o in host_off]
    pinned = [torch.from_numpy(o).pin_memory() for o in host_off]
    lap_pin = torch.empty(B, dtype=torch.float64).pin_memory()
    st_pin = torch.empty(B, dtype=torch.int32).pin_memory()
    ld = d_off[0].shape[1]
    lap = torch.empty(ld, dtype=torch.float64, device=dev)
    st = torch.empty(ld, dtype=torch.int32, device=dev)
    from spline_trajectory_optimization_b200 import sharding

    def step(j):
        l, s = ev.lap_times(d_off[j & 1], B=B, out=lap, status=st)
        best, idx = ev.argmin(l, s)
        # the path's only exchange: one (best lap, global candidate index) pair per rank, then argmin again
        return sharding.global_argmin(best, idx, rank * B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for j in range(args.warmup):
        step(j)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(args.steps):
        best = step(j)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = B * world * args.steps / (ms_total * 1e-3)
    # the clock sampler covers the device-timed region only: nvidia-smi polling takes the driver lock and would
    # otherwise stretch every synchronous host call of the e2e leg below
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers through the C ABI (H2D of the offsets + transpose + D2H of laps inside the call)
    for j in range(min(args.warmup, 2)):
        ev.lap_times_host_into(pinned[j & 1].data_ptr(), B, lap_pin.data_ptr(), st_pin.data_ptr())
    barrier()
    t0 = time.perf_counter()
    e2e_calls = []
    for j in range(args.steps):
        tc = time.perf_counter()
        ev.lap_times_host_into(pinned[j & 1].data_ptr(), B, lap_pin.data_ptr(), st_pin.data_ptr())
        _ = float(lap_pin.min())    # the step's result is consumed on the host
        e2e_calls.append(1e3 * (time.perf_counter() - tc))
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / float(e2e_s.item())

    # ---- roofline leg: per-kernel durations from CUDA events recorded on the launching stream inside the library
    lib.sto_set_stage_timing(1)
    stage = np.zeros((args.steps, 4), dtype=np.float32)
    for j in range(args.steps):
        ev.lap_times(d_off[j & 1], B=B, out=lap, status=st)
        buf = (ctypes.c_float * 4)()
        _lib.check(lib.sto_last_stage_ms(buf))
        stage[j] = list(buf)
    # informational: the same step with the lane-group block / cyclic-reduction fit solver (coefficients equal to
    # FITPACK's up to rounding instead of bit for bit)
    _lib.check(lib.sto_set_fit_solver(_lib.FIT_BLOCKS))
    stage_b = np.zeros((min(args.steps, 4), 4), dtype=np.float32)
    for j in range(stage_b.shape[0]):
        ev.lap_times(d_off[j & 1], B=B, out=lap, status=st)
        buf = (ctypes.c_float * 4)()
        _lib.check(lib.sto_last_stage_ms(buf))
        stage_b[j] = list(buf)
    _lib.check(lib.sto_set_fit_solver(_lib.FIT_FITPACK))
    ev.lap_times(d_off[(args.steps - 1) & 1], B=B, out=lap, status=st)   # `lap` again holds the default solver's result
    lib.sto_set_stage_timing(0)
    fp64_peak = ctypes.c_double(0.0)
    _lib.check(lib.sto_measure_fp64_peak(ctypes.byref(fp64_peak)))
    stage_ms = stage.mean(axis=0)
    qss_ms = float(stage_ms[3])
    peak, peak_src = measured_peaks()
    alg_bytes = (8 * M + 8) * B                      # offsets in, lap out, per candidate
    achieved = alg_bytes / (qss_ms * 1e-3) / 1e9
    ok = bool((st[:B] == 0).all().item())
    lap_host = lap[:B].cpu().numpy()

    # ---- informational: throughput at the per-GPU batch of BASELINE configs[2] (262,144 candidates / 8 GPUs)
    large = None
    if args.large_batch and rank == 0 and world == 1:
        BL = args.large_batch
        off_l = torch.from_numpy(np.tile(host_off[0], ((BL + B - 1) // B, 1))[:BL]).to(dev)
        d_l = ev.to_sample_major(off_l)
        del off_l
        ev.lap_times(d_l, B=BL)
        torch.cuda.synchronize()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(2):
            l_l, s_l = ev.lap_times(d_l, B=BL)
        t1e.record()
        torch.cuda.synchronize()
        large = {"candidates_per_step": BL, "value": 2 * BL / (t0e.elapsed_time(t1e) * 1e-3), "unit": UNIT,
                 "ms_per_step": t0e.elapsed_time(t1e) / 2, "all_status_ok": bool((s_l == 0).all().item()),
                 "laps_equal_small_batch": bool(torch.equal(l_l[:B].cpu(), torch.from_numpy(
                     ev.lap_times(d_off[0], B=B)[0].cpu().numpy()))),
                 "note": "same candidates tiled; bit planes move to global memory at this size (DESIGN.md section 2)"}
        del d_l

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: Monza, 4,096 lateral-offset candidates/GPU at 2 m "
                                   "(M=N=2895), FP64, exact reference QSS schedule",
                       "candidates_per_gpu": B, "M": M, "N": N, "qss_impl": args.qss,
                       "fit_solver": "fitpack: FITPACK's fpclos Givens sweep restated bit for bit (coefficients identical "
                                     "to the reference's scipy splprep)",
                       "l2": "two alternating candidate batches; one step touches ~1.5 GB (> 126 MB L2)",
                       "parallelism": f"candidate-sharded x{world}; all-gather of (lap, index) + argmin"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(8 * M * B),
                    "d2h_bytes_per_step": int(12 * B),
                    "api": "sto_lap_time_host_f64 (pinned host offsets [B][M] in, lap/status out)",
                    "ms_per_call": [round(x, 2) for x in e2e_calls]},
            "gpu_launches": int(5 * args.steps),
            "kernels_per_step": ["zero_status_kernel", "fit_kernel", "eval_kernel",
                                 "qss_memo_kernel" if args.qss == "memo" else "qss_plain_kernel", "argmin_kernel"],
            "roofline": {"bound": "hbm", "kernel": "qss_%s_kernel" % args.qss, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one qss_memo_kernel<8> launch at this
                         # workload, from profiles/r01_qss_memo_s5_ncu_summary.txt (ncu --set full); not live
                         "traffic": (4.290e9 if (args.qss == "memo" and B == CANDIDATES_PER_GPU) else None),
                         "traffic_source": "ncu --set full capture, profiles/r01_qss_memo_s5_ncu_summary.txt",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_candidate": 8 * M + 8, "kernel_ms": qss_ms,
                         "stage_ms": {"status": float(stage_ms[0]), "fit": float(stage_ms[1]),
                                      "sample": float(stage_ms[2]), "qss": qss_ms},
                         "note": "exact-schedule QSS is FP64-latency bound (~1e3 flop/B), not HBM bound; see DESIGN.md"},
            # the binding roof of the exact-schedule path is the FP64 pipe / its latency, not HBM.
            # "reference_schedule" counts the front steps the reference executes (70 flop each, ~433 k per line);
            # "executed" counts the evaluations the memoised kernel actually runs (~24.6 k per line, bit-identical result)
            "fp64": {"peak_tflops": fp64_peak.value, "peak_source": "measured live: sto_measure_fp64_peak (DFMA, 2 flop)",
                     "flops_per_candidate": {"fit": 60 * M, "sample": 200 * N, "qss_reference_schedule": 70 * 432793,
                                             "qss_executed": 70 * 24618},
                     "achieved_tflops_reference_schedule": (60 * M + 200 * N + 70 * 432793) * B / (float(stage_ms.sum()) * 1e-3) / 1e12,
                     "achieved_tflops_executed": (60 * M + 200 * N + 70 * 24618) * B / (float(stage_ms.sum()) * 1e-3) / 1e12,
                     "frac_executed": (60 * M + 200 * N + 70 * 24618) * B / (float(stage_ms.sum()) * 1e-3) / 1e12 / max(fp64_peak.value, 1e-9),
                     "note": "front-step / evaluation counts of the centre line (tests/hostsim counters); the path is "
                             "bound by the per-line dependent chain of FP64 divisions and square roots, see DESIGN.md"},
            "fast_fit_solver": {"solver": "blocks (lane groups: 32-block elimination + PCR over warp shuffles)",
                                "stage_ms": {"fit": float(stage_b.mean(axis=0)[1]), "sample": float(stage_b.mean(axis=0)[2]),
                                             "qss": float(stage_b.mean(axis=0)[3])},
                                "value": B * world / (float(stage_b.mean(axis=0).sum()) * 1e-3), "unit": UNIT,
                                "note": "coefficients within 2e-15 of FITPACK's; 1-2 lines in 500 then differ from the "
                                        "reference's lap by 1e-6..1e-4 s (schedule bifurcations), hence not the default"},
            "clocks": clocks,
            "lap_min_s": float(np.min(lap_host)), "lap_centre_line_s": float(lap_host[0]), "all_status_ok": ok,
        }
        if large is not None:
            line["large_batch"] = large
        if not args.no_cpu_baseline and world == 1:
            cb, olap = cpu_baseline_leg(rt, veh, host_off[(args.steps - 1) & 1])
            line["cpu_baseline"] = cb
            line["cpu_baseline"]["max_abs_lap_diff_vs_gpu_s"] = float(np.max(np.abs(olap - lap_host[:len(olap)])))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qss", default="memo", choices=["memo", "plain"])
    ap.add_argument("--candidates", type=int, default=CANDIDATES_PER_GPU, help="candidates per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--large-batch", type=int, default=32768,
                    help="also report device-resident throughput at this batch size (0 = skip; N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
