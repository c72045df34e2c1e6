#!/usr/bin/env python
"""bench.py -- candidate racing lines / second (fit + curvature + QSS lap time).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C]            # our arm (CUDA, libsto_b200.so)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]        # CPU arm: the oracle port, all host threads
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...            # one rank per GPU

A step = one pass of the hot path over one batch of synthetic candidates:
    offsets[M, B] (resident in HBM) -> periodic cubic fit -> resample + turn radius -> exact-schedule QSS -> lap[B]
    -> the path's only exchange: (config 1) local argmin, NCCL all-gather of one 16-byte (lap, index) pair per rank,
       argmin;  (configs 2-4) NCCL all-gather of the lap vector (8 B per candidate) + argmin on every rank.
Workloads (BASELINE.json `configs`; --config, default = 1 on one GPU, 2 on several):
    1  Monza at 2 m (M = N = 2895), 4,096 candidates per GPU (weak scaling)           <- the N = 1 headline
    2  Monza at 2 m, 262,144 candidates sharded over the GPUs (strong scaling)         <- the N > 1 default
    3  synthetic banked oval (800 m straights, R = 250 m, 9 deg bank), 65,536 candidates sharded
    4  Monza at 0.25 m (M = N = 23,160), --candidates lines sharded (the long-track path)
Prints ONE JSON line on rank 0 (see the task contract for the keys).
"""
import argparse
import ctypes
import glob
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CANDIDATES_PER_GPU = 4096
UNIT = "candidates/s"
METRIC = "candidate racing lines/sec (fit+curvature+QSS lap time), Monza"

WORKLOADS = {
    1: dict(track="monza", interval=2.0, per_gpu=CANDIDATES_PER_GPU, total=None, scaling="weak", reduce="pair",
            text="BASELINE configs[1]: Monza, 4,096 lateral-offset candidates/GPU at 2 m (M=N=2895), FP64, exact "
                 "reference QSS schedule"),
    2: dict(track="monza", interval=2.0, per_gpu=None, total=262144, scaling="strong", reduce="allgather",
            text="BASELINE configs[2]: Monza at 2 m (M=N=2895), 262,144 candidates sharded over the GPUs, all-gather "
                 "of the lap vector + argmin, FP64, exact reference QSS schedule"),
    3: dict(track="oval", interval=2.0, per_gpu=None, total=65536, scaling="strong", reduce="allgather",
            text="BASELINE configs[3]: synthetic banked oval (2 x 800 m straights, 2 x R=250 m, width 15 m, bank 0->9 deg "
                 "over 100 m; 4-column bounds), bank-aware QSS, 65,536 candidates sharded over the GPUs, FP64"),
    4: dict(track="monza", interval=0.25, per_gpu=None, total=16384, scaling="strong", reduce="allgather",
            text="BASELINE configs[4]: Monza at 0.25 m (M=N=23,160), candidates sharded over the GPUs, long-track fit + "
                 "QSS, FP64 (the full 1,048,576-line job = total/value seconds at this rate)"),
}


# ----------------------------------------------------------------------------------------------------------------
def build_track(interval=2.0, kind="monza", device=True):
    """Track as RaceTrack builds it (splines s=10 / k=3, samples every `interval` metres, bound look-up); device=False
    keeps everything on the host (no CUDA library is loaded: the reference arm)."""
    from spline_trajectory_optimization_b200 import candidates, tracks
    from spline_trajectory_optimization_b200.models.race_track import RaceTrack
    if kind == "oval":   # SURVEY.md 8(d), config 4 of the survey = BASELINE configs[3]
        centre, left, right = candidates.banked_oval(straight=800.0, radius=250.0, width=15.0, bank_deg=9.0, blend=100.0,
                                                     spacing=8.0)
        return RaceTrack("oval", left, right, centre, s=1.0, interval=interval, device=device)
    c, l, r = tracks.monza_raw()
    return RaceTrack("monza", l, r, c, s=10.0, interval=interval, device=device)


def track_bank(rt):
    from spline_trajectory_optimization_b200.models.trajectory import Trajectory
    bank = rt.center_d[:, Trajectory.BANK]
    return bank if np.any(bank != 0.0) else None


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


def kernel_source_sha16():
    """Identity of the kernels that produced a profile: sha256 over csrc/*.  profiles/*_traffic.json files carry the sha of
    the sources they were captured from; a figure whose sha does not match the sources in this tree is stale -> null."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "spline_trajectory_optimization_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        with open(os.path.join(csrc, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture,
    only if that capture was taken from the kernels in this tree (sha match); else None."""
    sha = kernel_source_sha16()
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                rec = json.load(f)
        except (OSError, ValueError):
            continue
        if rec.get("csrc_sha16") == sha and kernel in rec.get("kernels", {}):
            return float(rec["kernels"][kernel]["dram_bytes_per_launch"]), os.path.relpath(path, ROOT)
    return None, f"no capture of the current kernels (csrc sha {sha}) under profiles/"


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
def oracle_setup(veh):
    """The checker / CPU baseline (oracle/): the one place outside tests/ and smoke() that may execute it."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    O.build()
    p = veh.param
    ov = O.make_vehicle((p.max_lon_acc_mpss, p.max_lon_dcc_mpss, p.max_left_acc_mpss, p.max_right_acc_mpss,
                         p.max_speed_mps, p.max_jerk), veh.acc_intp.x, veh.acc_intp.c, veh.dcc_intp.x, veh.dcc_intp.c)
    return O, ov


def oracle_laps(O, ov, rt, offsets, ref_pow, threads):
    c, nrm = rt.center_d, rt.left_normals()
    bank = track_bank(rt)
    sb = np.zeros(len(c)) if bank is None else np.sin(bank)
    return O.lap_batch(c[:, 0], c[:, 1], nrm[:, 0], nrm[:, 1], offsets, c.ts(), sb, ov, n_threads=threads, ref_pow=ref_pow)


def cpu_baseline_leg(O, ov, rt, offsets, seconds_target=12.0):
    """The oracle port (plain C restatement of the reference's schedule, ref_pow=1 arithmetic) on all host threads over a
    bounded sample of the same candidates.  A reported baseline, not a target."""
    cores = os.cpu_count() or 1
    n0 = min(len(offsets), cores)
    t0 = time.perf_counter()
    oracle_laps(O, ov, rt, offsets[:n0], 1, cores)
    per_round = max(time.perf_counter() - t0, 1e-3)
    n = int(min(len(offsets), max(n0, cores * max(1, int(seconds_target / per_round)))))
    t0 = time.perf_counter()
    lap, st = oracle_laps(O, ov, rt, offsets[:n], 1, cores)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} of the step's candidates, oracle/sto_oracle.c lap_batch on {cores} threads, {dt:.2f} s"}, lap[:n]


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Python (~15 s per
    candidate at N = 2895); the arm times its compiled restatement (oracle/, pinned bit for bit on goldens of the
    unmodified reference) on all host threads, each step a bounded sample of the same workload.  Nothing of the product's
    CUDA library is loaded here: the track is built on the host (device=False)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = args.config or (1 if args.gpus <= 1 else 2)
    W = WORKLOADS[cfg]
    rt, veh = build_track(W["interval"], W["track"], device=False), test_vehicle()
    O, ov = oracle_setup(veh)
    cores = os.cpu_count() or 1
    M = len(rt.center_d)
    per_step = max(cores * 8, 64) if M < 10000 else cores      # ~35 ms / ~2.5 s per line per core
    off = make_offsets(rt, per_step, seed=1234)
    run = lambda: oracle_laps(O, ov, rt, off, 1, cores)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lap, st = run()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {"metric": METRIC if W["track"] == "monza" else METRIC.replace("Monza", "synthetic banked oval"),
            "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": W["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": W["text"], "config_id": cfg, "M": M, "N": M, "candidates_per_step": per_step,
                       "note": "bounded sample of the workload's candidates (throughput metric)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per_step} candidates/step on {cores} threads (oracle/sto_oracle.c)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "lap_min_s": float(np.nanmin(lap)), "all_status_ok": bool(not st.any())}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------------
def nccl_log_setup(rank):
    """NCCL's init report goes to a per-process file (stdout must carry exactly one JSON line); rank 0 echoes the lines
    that show the communicator (rank / nranks, NVLS) on stderr and counts them into the JSON line."""
    if "NCCL_DEBUG_FILE" in os.environ:
        return None
    d = os.path.join(tempfile.gettempdir(), "sto_nccl_%d" % os.getppid())
    os.makedirs(d, exist_ok=True)
    os.environ.setdefault("NCCL_DEBUG", "INFO")
    if os.environ["NCCL_DEBUG"].upper() in ("VERSION", "WARN"):
        os.environ["NCCL_DEBUG"] = "INFO"
    os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    os.environ["NCCL_DEBUG_FILE"] = os.path.join(d, "rank%d.%%p.log" % rank)
    return d


def nccl_log_report(d, world):
    if d is None:
        return {"log": "NCCL_DEBUG_FILE set by the caller"}
    ranks, nvls, lines = set(), False, []
    for path in sorted(glob.glob(os.path.join(d, "rank*.log"))):
        try:
            text = open(path, errors="replace").read().splitlines()
        except OSError:
            continue
        for ln in text:
            if "nranks" in ln and ("Init" in ln or "comm" in ln):
                lines.append(ln.strip())
                m = re.search(r"rank (\d+) nranks (\d+)", ln)
                if m and int(m.group(2)) == world:
                    ranks.add(int(m.group(1)))
            if "NVLS" in ln:
                nvls = True
    for ln in lines[:2 * world]:
        print(ln, file=sys.stderr)
    return {"ranks_seen": len(ranks), "nranks": world, "nvls_mentioned": nvls, "log_dir": d}


def fast_mode_leg(ev, d_small, B, exact_lap, rt, peak, big=131072):
    """The separately-reported FAST MODE (sto_lap_time_fast_f64: Thomas fit + sampler + two-sweep QSS + fill_time in one
    kernel): throughput lap-only and with every output column materialised, at the headline batch and at a large one;
    achieved HBM GB/s on ALGORITHMIC bytes against the measured peak; its measured lap error against the exact schedule
    on the same lines; and the track-table staging (bulk copy into shared memory) against warp-uniform global loads."""
    import torch
    from spline_trajectory_optimization_b200 import candidates
    M, N, dev = ev.M, ev.N, ev.device
    alg_lap = 8 * M + 8                                     # offsets in, lap out
    alg_full = 8 * M + 8 + 16 * (M + 3) + 64 * N            # + coefficients + 8 sample columns

    def timed(off, nb, reps, **kw):
        r = ev.lap_times_fast(off, B=nb, **kw)
        if kw.get("outputs"):
            kw["outputs"] = r[2]               # the timed calls write into the buffers of the warm-up call
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):                  # every repetition timed on its own; the fastest one is reported
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = ev.lap_times_fast(off, B=nb, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms_ = e0.elapsed_time(e1)
            best = ms_ if best is None or ms_ < best else best
        return best, r

    out = {"api": "sto_lap_time_fast_f64", "rounds": 2,
           "algorithmic_bytes_per_candidate": {"lap_only": alg_lap, "full_outputs": alg_full},
           "note": "NOT the reference's result: two sweeps with the reference's step operator instead of its lock-step "
                   "multi-front schedule; reported separately, never the default"}
    ms, r = timed(d_small, B, 5, rounds=2)
    fast_lap = r[0].cpu().numpy()
    err = fast_lap - exact_lap
    order_f, order_e = np.argsort(fast_lap), np.argsort(exact_lap)
    rank_f = np.empty(B); rank_f[order_f] = np.arange(B)
    rank_e = np.empty(B); rank_e[order_e] = np.arange(B)
    out["error_vs_exact_schedule"] = {
        "candidates": int(B), "max_abs_lap_err_s": float(np.max(np.abs(err))), "mean_lap_err_s": float(np.mean(err)),
        "std_lap_err_s": float(np.std(err)), "spearman_rank_correlation": float(np.corrcoef(rank_f, rank_e)[0, 1]),
        "exact_best_is_in_fast_top": int(rank_f[order_e[0]]) + 1}
    out["max_abs_lap_err_s"] = out["error_vs_exact_schedule"]["max_abs_lap_err_s"]
    ms_g, _ = timed(d_small, B, 5, rounds=2, stage_tables=0)
    ms_f, _ = timed(d_small, B, 3, rounds=2, outputs=True)
    out["batch_%d" % B] = {"ms_lap_only": ms, "value_lap_only": B / (ms * 1e-3), "ms_full_outputs": ms_f,
                           "value_full_outputs": B / (ms_f * 1e-3), "unit": UNIT,
                           "ms_lap_only_tables_in_global_memory": ms_g}
    free_b, _tot = torch.cuda.mem_get_info(dev)
    per_cand = ev.lib.sto_fast_workspace_bytes(M, N, 1024) / 1024.0 + 8.0 * (2 * (M + 3) + 8 * N) + 8.0 * M
    BL = int(min(big, (int(0.6 * free_b / per_cand) // 32) * 32))
    d_l = candidates.smooth_offsets_device(M, 0, BL, rt.dist_to_left, rt.dist_to_right, dev, seed=77)
    ms_ls, r_l = timed(d_l, BL, 3, rounds=2, stage_tables=1)
    ms_lg, _ = timed(d_l, BL, 3, rounds=2, stage_tables=0)
    ms_l, _ = timed(d_l, BL, 3, rounds=2)                      # the library's own choice
    ms_lf, r_lf = timed(d_l, BL, 3, rounds=2, outputs=True)
    ok = bool((r_l[1] == 0).all().item()) and bool(torch.equal(r_l[0], r_lf[0]))
    del d_l, r_lf
    out["batch_%d" % BL] = {"ms_lap_only": ms_l, "value_lap_only": BL / (ms_l * 1e-3), "ms_full_outputs": ms_lf,
                            "value_full_outputs": BL / (ms_lf * 1e-3), "unit": UNIT,
                            "ms_lap_only_tables_in_global_memory": ms_lg, "ms_lap_only_tables_staged": ms_ls,
                            "all_status_ok_and_laps_equal": ok}
    ach_lap = alg_lap * BL / (ms_l * 1e-3) / 1e9
    ach_full = alg_full * BL / (ms_lf * 1e-3) / 1e9
    out["roofline"] = {"bound": "hbm", "kernel": "fast_kernel", "peak": peak, "unit": "GB/s",
                       "achieved": ach_full, "frac": ach_full / peak, "counted_on": "full outputs, batch %d" % BL,
                       "achieved_lap_only": ach_lap, "frac_lap_only": ach_lap / peak,
                       "traffic": ncu_traffic("fast_kernel")[0], "traffic_source": ncu_traffic("fast_kernel")[1]
                       + " (full outputs, 131,072 candidates per launch)",
                       "note": "lap-only the path is ~70 flop per algorithmic byte (FP64-pipe side of the 5.7 flop/B "
                               "balance point); with the outputs materialised ~6 flop/B"}
    out["table_staging"] = {"how": "cp.async.bulk (UBLKCP) + mbarrier, six tables once per CTA into shared memory",
                            "speedup_small_batch": ms_g / ms, "speedup_large_batch": ms_lg / ms_ls,
                            "policy": "staged while the launch is a single wave (B <= 148 x 512), global loads beyond"}
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from spline_trajectory_optimization_b200 import _lib, candidates, sharding
    from spline_trajectory_optimization_b200.evaluator import BatchedLineEvaluator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU arm has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl_dir = None
    if world > 1:
        nccl_dir = nccl_log_setup(rank)
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    cfg = args.config or (1 if world == 1 else 2)
    W = WORKLOADS[cfg]
    strong = W["scaling"] == "strong"

    rt, veh = build_track(W["interval"], W["track"]), test_vehicle()
    M = N = len(rt.center_d)
    bank = track_bank(rt)
    ev = BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh, bank=bank, device=dev,
                              impl=args.qss)
    if strong:
        total = int(args.candidates or W["total"])
        lo, hi = sharding.shard_range(total, rank, world)
    else:
        per_gpu = int(args.candidates or W["per_gpu"])
        total = per_gpu * world
        lo, hi = rank * per_gpu, (rank + 1) * per_gpu
    B = hi - lo                                                   # this rank's candidates per step
    # launches of at most `chunk` candidates: the per-candidate workspace (~140 B per sample) bounds a launch
    per_cand = lib.sto_lap_workspace_bytes(M, N, 1024, ev.impl) / 1024.0
    chunk = int(min(B, 131072, max(32, (int(args.launch_gb * 2 ** 30 / per_cand) // 32) * 32)))
    cuts = [(c0, min(B, c0 + chunk)) for c0 in range(0, B, chunk)]
    # Inputs: config 1 keeps round 1's NumPy-generated lines (two alternating batches per rank); the sharded configs
    # generate theirs on the device (candidate b depends on (seed, b) only, not on the sharding).  Two alternating input
    # sets whenever one set is smaller than ~8x the 126 MB L2, so no step re-reads what its predecessor left there.
    n_sets = 2 if (8.0 * M * B < 1.0e9) else 1
    d_off, host_first = [], None
    for j in range(n_sets):
        if not strong:
            ho = make_offsets(rt, B, seed=1234 + 7919 * rank + 104729 * j)
            if j == 0:
                host_first = ho
            d_off.append([ev.to_sample_major(torch.from_numpy(ho[a:b]).to(dev)) for a, b in cuts])
        else:
            d_off.append([candidates.smooth_offsets_device(M, lo + a, lo + b, rt.dist_to_left, rt.dist_to_right, dev,
                                                           seed=1234 + 104729 * j) for a, b in cuts])
    # Steps are software-pipelined over two CUDA streams: step k+1's fit / sampler / QSS start while the slowest warps of
    # step k's QSS are still running (a QSS launch ends with its hardest candidates; the SMs the others leave are idle
    # otherwise).  Each stream has its own workspace, outputs and reduction buffers; every step is a complete pass.  The
    # strictly serial figure (one step at a time on one stream) is reported next to it (`serial`).
    n_slots = args.slots if (args.pipeline and args.slots * per_cand * chunk <= 40e9) else 1

    class Slot:
        def __init__(self, k):
            self.ev = ev if k == 0 else BatchedLineEvaluator(rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts(), veh,
                                                             bank=bank, device=dev, impl=args.qss)
            self.stream = torch.cuda.Stream(device=dev) if n_slots > 1 else torch.cuda.current_stream(dev)
            self.lap = torch.empty(B, dtype=torch.float64, device=dev)
            self.st = torch.empty(B, dtype=torch.int32, device=dev)
            self.lap_ld = [torch.empty(((b - a + 31) & ~31), dtype=torch.float64, device=dev) for a, b in cuts]
            self.st_ld = [torch.empty(((b - a + 31) & ~31), dtype=torch.int32, device=dev) for a, b in cuts]
            self.pair_argmin = sharding.DeviceArgmin(dev)
            self.best = torch.empty(1, dtype=torch.float64, device=dev)
            self.best_idx = torch.empty(1, dtype=torch.int64, device=dev)

    slots = [Slot(k) for k in range(n_slots)]
    lap_ld, st_ld = slots[0].lap_ld, slots[0].st_ld

    def evaluate(j, S=None):
        S = S or slots[0]
        for k, (a, b) in enumerate(cuts):
            l, s = S.ev.lap_times(d_off[j % n_sets][k], B=b - a, out=S.lap_ld[k], status=S.st_ld[k])
            if len(cuts) > 1:
                S.lap[a:b].copy_(l)
                S.st[a:b].copy_(s)
        return (S.lap, S.st) if len(cuts) > 1 else (S.lap_ld[0][:B], S.st_ld[0][:B])

    def reduce(l, s, S=None):
        S = S or slots[0]
        if W["reduce"] == "pair":      # one (best lap, global index) pair per rank: 16-byte all-gather, argmin
            return S.pair_argmin(l, s, lo)
        # the lap vector of the whole job on every rank (8 B per candidate over NVLink), argmin there; failed candidates
        # travel as NaN, which the argmin ignores
        full = sharding.all_gather_laps(torch.where(s == 0, l, torch.full_like(l, float("nan"))), total)
        _lib.check(lib.sto_argmin_f64(ctypes.c_void_p(full.data_ptr()), None, total, ctypes.c_void_p(S.best.data_ptr()),
                                      ctypes.c_void_p(S.best_idx.data_ptr()),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return S.best, S.best_idx

    def step(j, S=None):
        return reduce(*evaluate(j, S), S)

    def run_steps(n, pipelined):
        """n complete steps; returns (elapsed ms on the device, the last step's winner)."""
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        t0e.record(cur)
        win = None
        if pipelined and n_slots > 1:
            for S in slots:
                S.stream.wait_event(t0e)
            for j in range(n):
                S = slots[j % n_slots]
                with torch.cuda.stream(S.stream):
                    win = step(j, S)
            for S in slots:
                cur.wait_stream(S.stream)
        else:
            for j in range(n):
                win = step(j)
        t1e.record(cur)
        return t0e, t1e, win

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(args.warmup, True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1, winner = run_steps(args.steps, True)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = total * args.steps / (ms_total * 1e-3)
    # the clock sampler covers the device-timed region only: nvidia-smi polling takes the driver lock and would
    # otherwise stretch every synchronous host call of the e2e leg below
    clocks = sampler.stop() if rank == 0 else None
    win_lap, win_idx = float(winner[0].item()), int(winner[1].item())
    serial = None
    if n_slots > 1:   # the same steps strictly one at a time
        n_ser = min(args.steps, 6)
        run_steps(2, False)
        barrier()
        s0, s1, _w = run_steps(n_ser, False)
        barrier()
        ms_s = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_s, op=dist.ReduceOp.MAX)
        serial = {"steps": n_ser, "ms_per_step": float(ms_s.item()) / n_ser,
                  "value": total * n_ser / (float(ms_s.item()) * 1e-3), "unit": UNIT,
                  "note": "one step at a time on one stream (no overlap between consecutive steps)"}

    # ---- e2e: host buffers through the C ABI (H2D of the offsets + transpose + D2H of laps inside the call), then the
    # same exchange as above starting from the host laps
    pinned = []
    for j in range(n_sets):
        hp = torch.empty((B, M), dtype=torch.float64).pin_memory()
        for k, (a, b) in enumerate(cuts):
            hp[a:b].copy_(d_off[j][k][:, :b - a].T)
        pinned.append(hp)
    lap_pin = torch.empty(B, dtype=torch.float64).pin_memory()
    st_pin = torch.empty(B, dtype=torch.int32).pin_memory()
    lap_dev = torch.empty(B, dtype=torch.float64, device=dev)
    st_dev = torch.empty(B, dtype=torch.int32, device=dev)

    def e2e_step(j, bufs):
        lp, sp = bufs
        ev.lap_times_host_into(pinned[j % n_sets].data_ptr(), B, lp.data_ptr(), sp.data_ptr(),
                               int(args.launch_gb * 2 ** 30))
        if world > 1:
            lap_dev.copy_(lp, non_blocking=True)
            st_dev.copy_(sp, non_blocking=True)
            b_, i_ = reduce(lap_dev, st_dev)
            return float(b_.item())                           # the step's result is consumed on the host
        return float(np.nanmin(lp.numpy()))

    # Two host threads drive the GPU when the steps are pipelined (the library keeps two contexts per device: the calls
    # overlap on two streams); every call is the plain synchronous sto_lap_time_host_f64 on its own pinned buffers.
    n_thr = 2 if (n_slots > 1 and world == 1) else 1
    bufs = [(lap_pin, st_pin)] + [(torch.empty(B, dtype=torch.float64).pin_memory(),
                                   torch.empty(B, dtype=torch.int32).pin_memory()) for _ in range(n_thr - 1)]
    e2e_calls = []

    def e2e_run(n, record):
        def worker(t):
            for j in range(t, n, n_thr):
                tc = time.perf_counter()
                e2e_step(j, bufs[t])
                if record:
                    e2e_calls.append(1e3 * (time.perf_counter() - tc))
        if n_thr == 1:
            worker(0)
        else:
            th = [threading.Thread(target=worker, args=(t,)) for t in range(n_thr)]
            for x in th:
                x.start()
            for x in th:
                x.join()

    e2e_run(max(min(args.warmup, 2), n_thr), False)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps, True)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = total * args.steps / float(e2e_s.item())
    e2e_one = None
    if n_thr > 1:   # the same calls from one thread, back to back
        n1 = min(args.steps, 6)
        t1 = time.perf_counter()
        for j in range(n1):
            e2e_step(j, bufs[0])
        e2e_one = {"value": total * n1 / (time.perf_counter() - t1), "unit": UNIT, "calls": n1,
                   "note": "one host thread, one call at a time"}
    del pinned

    # ---- roofline leg: per-kernel durations from CUDA events recorded on the launching stream inside the library
    lib.sto_set_stage_timing(1)
    n_rf = min(args.steps, 4 if len(cuts) > 1 or B > 16384 else args.steps)
    stage = np.zeros((n_rf, 4), dtype=np.float64)
    for j in range(n_rf):
        for k, (a, b) in enumerate(cuts):
            ev.lap_times(d_off[j % n_sets][k], B=b - a, out=lap_ld[k], status=st_ld[k])
            buf = (ctypes.c_float * 4)()
            _lib.check(lib.sto_last_stage_ms(buf))
            stage[j] += np.array(list(buf))
    stage_ms = stage.mean(axis=0)
    fast_fit = None
    if cfg == 1:
        # informational: the same step with the lane-group block / cyclic-reduction fit solver (coefficients equal to
        # FITPACK's up to rounding instead of bit for bit)
        _lib.check(lib.sto_set_fit_solver(_lib.FIT_BLOCKS))
        stage_b = np.zeros((min(args.steps, 4), 4), dtype=np.float32)
        for j in range(stage_b.shape[0]):
            ev.lap_times(d_off[j % n_sets][0], B=B, out=lap_ld[0], status=st_ld[0])
            buf = (ctypes.c_float * 4)()
            _lib.check(lib.sto_last_stage_ms(buf))
            stage_b[j] = list(buf)
        _lib.check(lib.sto_set_fit_solver(_lib.FIT_FITPACK))
        sb_ = stage_b.mean(axis=0)
        fast_fit = {"solver": "blocks (lane groups: 32-block elimination + PCR over warp shuffles)",
                    "stage_ms": {"fit": float(sb_[1]), "sample": float(sb_[2]), "qss": float(sb_[3])},
                    "value": B * world / (float(sb_.sum()) * 1e-3), "unit": UNIT,
                    "note": "coefficients within 2e-15 of FITPACK's; 1-2 lines in 500 then differ from the reference's "
                            "lap by 1e-6..1e-4 s (schedule bifurcations), hence not the default"}
    lib.sto_set_stage_timing(0)
    l_fin, s_fin = evaluate(0)                                # laps of input set 0 with the default solver, for the checks
    torch.cuda.synchronize()
    lap_host, st_host = l_fin.cpu().numpy().copy(), s_fin.cpu().numpy().copy()
    ok = bool(not st_host.any())
    ok_t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    all_ok = bool(ok_t.item() == 1.0)

    qss_ms = float(stage_ms[3])
    peak, peak_src = measured_peaks()
    alg_bytes = (8 * M + 8) * B                      # offsets in, lap out, per candidate (this rank's launch set)
    achieved = alg_bytes / (qss_ms * 1e-3) / 1e9
    qss_kernel = lib.sto_last_qss_kernel().decode() or ("qss_%s_kernel" % args.qss)   # the kernel the launches above used
    traffic, traffic_src = ncu_traffic(qss_kernel) if (cfg == 1 and B == CANDIDATES_PER_GPU) else (None, "captured at config 1 only")

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- rank 0: live counts, FP64 roof, parity sample against the oracle, CPU baseline, the JSON line
    fp64_peak = ctypes.c_double(0.0)
    _lib.check(lib.sto_measure_fp64_peak(ctypes.byref(fp64_peak)))
    # front steps per line, LIVE: the first candidates of this run through the staged path with the summary rows
    n_cnt = int(min(B, 64 if M < 10000 else 8))
    d_cnt = d_off[0][0][:, :((n_cnt + 31) & ~31)].contiguous()
    u_, cx_, cy_, _s = ev.fit(d_cnt, B=n_cnt)
    S_ = ev.sample(u_, cx_, cy_, B=n_cnt, want=("x", "y", "radius"))
    q_ = ev.qss(S_["x"], S_["y"], S_["radius"], B=n_cnt, sin_bank=ev.sinb_host, profiles=False)
    steps_mean = float(q_["summary"][6, :n_cnt].mean().item())
    iters_mean = float(q_["summary"][7, :n_cnt].mean().item())
    flops_ref = 60 * M + 200 * N + 70 * steps_mean
    step_s = float(stage_ms.sum()) * 1e-3

    O, ov = oracle_setup(veh)
    cores = os.cpu_count() or 1
    n_par = int(min(B, args.parity or (64 if M < 10000 else max(2, min(8, cores // 4)))))
    off_par = d_off[0][0][:, :n_par].T.contiguous().cpu().numpy() if strong else host_first[:n_par]
    t0 = time.perf_counter()
    o0, os0 = oracle_laps(O, ov, rt, off_par, 0, cores)
    o1, os1 = oracle_laps(O, ov, rt, off_par, 1, cores)
    parity = {"candidates": n_par, "oracle_s": round(time.perf_counter() - t0, 2),
              "bit_exact_vs_oracle_product_arithmetic": bool(np.array_equal(lap_host[:n_par], o0)),
              "max_abs_lap_diff_vs_oracle_reference_arithmetic_s": float(np.max(np.abs(lap_host[:n_par] - o1))),
              "tolerance_s": 1e-6, "oracle_status_ok": bool(not os0.any() and not os1.any())}

    line = {
        "metric": METRIC if W["track"] == "monza" else METRIC.replace("Monza", "synthetic banked oval"),
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": W["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": W["text"], "config_id": cfg, "candidates_total": total, "candidates_per_gpu": B,
                   "candidates_per_launch": chunk, "launches_per_step": len(cuts), "M": M, "N": N,
                   "qss_impl": args.qss, "bank": bank is not None,
                   "pipelining": ("consecutive steps rotate over %d CUDA streams (own workspace each): step k+1 starts " % n_slots +
                                  "under the tail of step k's QSS; every step is a complete pass; serial figure in `serial`"
                                  if n_slots > 1 else "none (one stream)"),
                   "fit_solver": "fitpack: FITPACK's fpclos Givens sweep restated bit for bit (coefficients identical "
                                 "to the reference's scipy splprep)",
                   "l2": ("two alternating candidate sets" if n_sets == 2 else "one candidate set of %.1f GB (>> 126 MB L2)"
                          % (8e-9 * M * B)) + "; one step touches ~%.1f GB of state" % (per_cand * B / 1e9),
                   "parallelism": (f"candidate-sharded x{world}; all-gather of one (lap, index) pair per rank + argmin"
                                   if W["reduce"] == "pair" else
                                   f"candidate-sharded x{world}; NCCL all-gather of the lap vector ({8 * total} B) + argmin "
                                   "on every rank")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(8 * M * B + (12 * B if world > 1 else 0)),
                "d2h_bytes_per_step": int(12 * B + (16 if world > 1 else 0)),
                "host_threads": n_thr, "one_thread": e2e_one,
                "api": "sto_lap_time_host_f64 (pinned host offsets [B][M] in, lap/status out)"
                       + (", called from %d host threads (two contexts per device in the library: the calls overlap)" % n_thr
                          if n_thr > 1 else "")
                       + ("; then H2D of the laps, the same NCCL exchange, D2H of the winner" if world > 1 else ""),
                "ms_per_call": [round(x, 2) for x in e2e_calls]},
        "gpu_launches": int((4 * len(cuts) + (3 if W["reduce"] == "pair" else 1)) * args.steps),
        "kernels_per_step": ["zero_status_kernel", "fit_kernel", "eval_kernel", qss_kernel] +
                            (["argmin_kernel", "argmin_pair_kernel", "argmin_pairs_kernel"] if W["reduce"] == "pair"
                             else ["argmin_kernel"]),
        "roofline": {"bound": "hbm", "kernel": qss_kernel, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_bytes_per_candidate": 8 * M + 8, "kernel_ms": qss_ms,
                     "stage_ms": {"status": float(stage_ms[0]), "fit": float(stage_ms[1]),
                                  "sample": float(stage_ms[2]), "qss": qss_ms},
                     "note": "exact-schedule QSS is bound by per-line dependent FP64 chains (~1e3 flop/B), not by HBM; the "
                             "HBM-bound variant of the path is the `fast_mode` block; see DESIGN.md"},
        # the second roof.  Front steps per line are read LIVE from this run's summary rows.
        "fp64": {"peak_tflops": fp64_peak.value, "peak_source": "measured live: sto_measure_fp64_peak (DFMA, 2 flop)",
                 "front_steps_per_line_live": steps_mean, "outer_iterations_per_line_live": iters_mean,
                 "counted_on": f"the first {n_cnt} candidates of this run (summary rows of sto_qss_f64)",
                 "flops_per_candidate_reference_schedule": flops_ref,
                 "achieved_tflops_reference_schedule": flops_ref * B / step_s / 1e12,
                 "frac_reference_schedule": flops_ref * B / step_s / 1e12 / max(fp64_peak.value, 1e-9),
                 "note": "70 flop per front step of the reference's schedule; the memoised kernel evaluates ~6 % of them "
                         "(bit-identical result), so its executed FP64 work is ~17x lower: the path is bound by per-line "
                         "dependent chains of divisions / square roots, neither roof"},
        "serial": serial,
        "parity": parity,
        "winner": {"lap_s": win_lap, "candidate": win_idx},
        "clocks": clocks,
        "lap_min_s": float(np.nanmin(lap_host)), "lap_first_candidate_s": float(lap_host[0]), "all_status_ok": all_ok,
    }
    if fast_fit is not None:
        line["fast_fit_solver"] = fast_fit
    if world > 1:
        line["nccl"] = nccl_log_report(nccl_dir, world)
    # ---- informational (N = 1, config 1): device-resident throughput at BASELINE configs[2]'s per-GPU batch on 8 GPUs
    if args.large_batch and world == 1 and cfg == 1:
        BL = int(args.large_batch)
        d_l = candidates.smooth_offsets_device(M, 0, BL, rt.dist_to_left, rt.dist_to_right, dev, seed=99)
        ev.lap_times(d_l, B=BL)
        torch.cuda.synchronize()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(2):
            l_l, s_l = ev.lap_times(d_l, B=BL)
        t1e.record()
        torch.cuda.synchronize()
        line["large_batch"] = {"candidates_per_step": BL, "value": 2 * BL / (t0e.elapsed_time(t1e) * 1e-3), "unit": UNIT,
                               "ms_per_step": t0e.elapsed_time(t1e) / 2, "all_status_ok": bool((s_l == 0).all().item()),
                               "note": "262,144 / 8 candidates in one launch, one launch at a time; bit planes in global "
                                       "memory at this size (DESIGN.md section 2)"}
        if n_slots > 1:
            # the same launches with two in flight (one per stream, own workspace each), as the timed steps above run:
            # 32,768 candidates put 7 warps on an SM, a second launch fills the rest
            cur = torch.cuda.current_stream(dev)
            for S in slots[:2]:
                with torch.cuda.stream(S.stream):
                    S.ev.lap_times(d_l, B=BL)
            torch.cuda.synchronize()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record(cur)
            n_l = 4
            for S in slots[:2]:
                S.stream.wait_event(t0e)
            for j in range(n_l):
                S = slots[j % 2]
                with torch.cuda.stream(S.stream):
                    l_l, s_l = S.ev.lap_times(d_l, B=BL)
            for S in slots[:2]:
                cur.wait_stream(S.stream)
            t1e.record(cur)
            torch.cuda.synchronize()
            line["large_batch"]["pipelined"] = {"value": n_l * BL / (t0e.elapsed_time(t1e) * 1e-3), "unit": UNIT,
                                                "ms_per_step": t0e.elapsed_time(t1e) / n_l, "launches_in_flight": 2,
                                                "all_status_ok": bool((s_l == 0).all().item())}
        del d_l
    if args.large_batch and world == 1 and cfg == 1:
        # ---- informational: BASELINE configs[2] (262,144 candidates) on ONE GPU, the strong-scaling reference of the
        # multi-GPU lines (bench.py --gpus N runs that configuration sharded): two launches of 131,072
        tot2, ch2 = WORKLOADS[2]["total"], 131072
        t_ms = 0.0
        ok2 = True
        for rep in range(2):                       # first pass warms up
            t_ms = 0.0
            for c0 in range(0, tot2, ch2):
                d_c = candidates.smooth_offsets_device(M, c0, c0 + ch2, rt.dist_to_left, rt.dist_to_right, dev, seed=1234)
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                l_c, s_c = ev.lap_times(d_c, B=ch2)
                a1.record()
                torch.cuda.synchronize()
                t_ms += a0.elapsed_time(a1)
                ok2 = ok2 and bool((s_c == 0).all().item())
                del d_c
        line["config2_on_one_gpu"] = {"candidates": tot2, "launches": tot2 // ch2, "ms": t_ms, "value": tot2 / (t_ms * 1e-3),
                                      "unit": UNIT, "all_status_ok": ok2,
                                      "note": "the workload `bench.py --gpus N` (N > 1) shards; device time of the launches only"}
    if world == 1 and cfg == 1 and not args.no_fast_mode:
        line["fast_mode"] = fast_mode_leg(ev, d_off[0][0], B, lap_host, rt, peak)
    if not args.no_cpu_baseline and world == 1:
        base_off = host_first if not strong else d_off[0][0][:, :min(B, 4096)].T.contiguous().cpu().numpy()
        cb, olap = cpu_baseline_leg(O, ov, rt, base_off)
        line["cpu_baseline"] = cb
        line["cpu_baseline"]["max_abs_lap_diff_vs_gpu_s"] = float(np.max(np.abs(olap - lap_host[:len(olap)])))
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4],
                    help="BASELINE.json configs[] index (0 = 1 on one GPU, 2 on several)")
    ap.add_argument("--qss", default="memo", choices=["memo", "plain"])
    ap.add_argument("--candidates", type=int, default=0,
                    help="candidates per GPU (config 1) / in total (configs 2-4); 0 = the configuration's own size")
    ap.add_argument("--launch-gb", type=float, default=56.0, help="device workspace budget of one launch")
    ap.add_argument("--parity", type=int, default=0, help="candidates checked against the oracle (0 = choose)")
    ap.add_argument("--slots", type=int, default=2, help="streams the steps are pipelined over")
    ap.add_argument("--no-pipeline", dest="pipeline", action="store_false",
                    help="run the timed steps strictly one at a time (default: two-stream software pipelining)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the separately-reported fast-mode block")
    ap.add_argument("--large-batch", type=int, default=32768,
                    help="config 1, N = 1: also report device-resident throughput at this batch size (0 = skip)")
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
