"""Developer tool (CPU): event log of the memoised QSS schedule (host build of the device code) -> where the rounds go.
Usage: python tools/analysis/qss_events.py [candidate index in tests/golden/cand_m2895_n2895.npz]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
from helpers import golden, veh_args  # noqa: E402
import hostsim_py as H  # noqa: E402

KIND = {1: "KILL", 2: "STOP", 3: "WRITE", 4: "KEEP", 5: "SPAWN", 6: "RESPAWN", 7: "ZERO"}


def events(case="cand_m2895_n2895", idx=0, impl=1):
    d = golden(case)
    veh = H.make_vehicle(*veh_args(d))
    L = H.lib()
    L.hostsim_log_size.restype = C.c_longlong
    L.hostsim_log_enable(1)
    r = H.qss(impl, d["ref_X"][idx:idx + 1], d["ref_Y"][idx:idx + 1], d["ref_CURVATURE"][idx:idx + 1], None, veh)
    n = L.hostsim_log_size()
    buf = np.empty(n, dtype=np.int32)
    L.hostsim_log_copy(buf.ctypes.data_as(C.POINTER(C.c_int)))
    L.hostsim_log_enable(0)
    return buf.reshape(-1, 4), r, d["ref_X"].shape[1]


if __name__ == "__main__":
    idx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    ev, r, N = events(idx=idx)
    print("events", len(ev), "lap", r["lap"][0], "iters", int(ev[:, 0].max()) + 1)
    for ph, name in enumerate(("B-orig", "B-spawn", "F-orig", "F-spawn")):
        e = ev[ev[:, 1] == ph]
        kinds = {KIND[k]: int((e[:, 3] == k).sum()) for k in np.unique(e[:, 3])}
        per_iter = np.bincount(e[:, 0], minlength=ev[:, 0].max() + 1)
        print(f"{name:8s} events {len(e):6d}  {kinds}  per-iteration max {per_iter.max()} mean {per_iter.mean():.1f} "
              f"iterations with any {int((per_iter > 0).sum())}")


def chain_stats(ev, N, ph, fwd, G=(1, 2, 4, 8, 32)):
    """Lower bound on evaluation rounds of one kind of sub-pass when independent dependency chains run in parallel.
    Dependency (exact): event j needs event i < j of the same sub-pass when i WROTE the sample j reads (its p), or
    they target the same sample."""
    e = ev[ev[:, 1] == ph]
    tot = {g: 0 for g in G}
    longest = 0
    for it in np.unique(e[:, 0]):
        x = e[e[:, 0] == it]
        depth = {}
        wrote = {}      # sample -> depth of the event that wrote it
        target = {}     # sample -> depth of last event targeting it
        dmax = 0
        for _, _, p, k in x:
            q = (p + 1) % N if fwd else (p - 1) % N
            d = 1 + max(wrote.get(p, 0), target.get(q, 0), wrote.get(q, 0))
            # an event whose source was targeted (read/written) later... reading p after someone wrote p handled above
            if k in (3, 5):
                wrote[q] = d
            target[q] = d
            dmax = max(dmax, d)
        for g in G:
            tot[g] += max(dmax, -(-len(x) // g))
        longest = max(longest, dmax)
    return len(e), tot, longest


if __name__ == "__main__":
    for ph, name, fwd in ((0, "B-orig", False), (1, "B-spawn", False), (2, "F-orig", True), (3, "F-spawn", True)):
        n, tot, longest = chain_stats(ev, N, ph, fwd)
        print(f"{name:8s} evals {n:6d}  rounds lower bound by lanes {tot}  longest chain {longest}")


def forig_run_schedule(ev, G=(1, 2, 4, 8, 16, 32)):
    """F-orig under the exact 'live-run' rule: evaluated fronts separated by a dead row are independent chains; a run of
    live rows is one sequential chain.  Rounds = greedy list scheduling of the chains of each sub-pass on G lanes."""
    e = ev[(ev[:, 1] == 10) | (ev[:, 1] == 11)]
    tot = {g: 0 for g in G}
    nchains = 0
    for it in np.unique(e[:, 0]):
        x = e[e[:, 0] == it]
        chains = []
        for _, ph, row, _ in x:
            if ph == 11 or not chains:
                chains.append(0)
            chains[-1] += 1
        nchains += len(chains)
        for g in G:
            lanes = [0] * g
            for c in chains:                       # in row order, next free lane
                k = int(np.argmin(lanes))
                lanes[k] += c
            tot[g] += max(lanes)
    return len(e), nchains, tot


if __name__ == "__main__":
    n, nch, tot = forig_run_schedule(ev)
    print(f"F-orig live-run rule: {n} evaluations in {nch} chains; rounds by lanes {tot}")
