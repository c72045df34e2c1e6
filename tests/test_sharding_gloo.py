"""CPU: the N > 1 host logic (candidate sharding, (lap, index) all-gather + argmin, full lap all-gather) on a
world_size-2 gloo group.  The per-shard "laps" come from the oracle on a small Monza batch."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import golden, veh_args


def _worker(rank, world, port, laps_all, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spline_trajectory_optimization_b200 import sharding
    B = len(laps_all)
    lo, hi = sharding.shard_range(B, rank, world)
    local = torch.from_numpy(laps_all[lo:hi].copy())
    valid = ~torch.isnan(local)
    if valid.any():
        key = torch.where(valid, local, torch.full_like(local, float("inf")))
        li = torch.argmin(key)
        best, idx = key[li], li
    else:
        best, idx = torch.tensor(float("nan"), dtype=torch.float64), torch.tensor(-1)
    g_best, g_idx = sharding.global_argmin(best, idx, lo)
    full = sharding.all_gather_laps(local, B)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[float(g_best), float(g_idx)], full.numpy()]))
    dist.destroy_process_group()


def _run(laps_all, tmp_path, world=2, port=29631):
    mp.spawn(_worker, args=(world, port, laps_all, str(tmp_path)), nprocs=world, join=True)
    return [np.load(os.path.join(tmp_path, f"r{r}.npy")) for r in range(world)]


def _oracle_laps():
    import oracle_py as O
    d = golden("cand_m579_n579")
    veh = O.make_vehicle(*veh_args(d))
    nx, ny = -np.sin(d["centre_yaw"]), np.cos(d["centre_yaw"])
    lap, st = O.lap_batch(d["centre_x"], d["centre_y"], nx, ny, d["offsets"][:7], d["ts"], np.zeros(len(d["ts"])), veh,
                          n_threads=2, ref_pow=0)
    assert not st.any()
    return lap


def test_shard_range_partitions():
    from spline_trajectory_optimization_b200.sharding import shard_range
    for B in (1, 7, 8, 4096, 262144, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(B, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_argmin_and_allgather_world2(tmp_path):
    laps = _oracle_laps()                       # 7 candidates: ragged over 2 ranks (3 + 4)
    res = _run(laps, tmp_path)
    for r in res:
        assert r[0] == laps.min() and int(r[1]) == int(np.argmin(laps))
        assert np.array_equal(r[2:], laps)
    assert np.array_equal(res[0], res[1])


@pytest.mark.timeout(300)
def test_argmin_ignores_nan_and_breaks_ties_low(tmp_path):
    laps = np.array([np.nan, 105.5, 104.25, np.nan, 104.25, 110.0, np.nan, np.nan])
    res = _run(laps, tmp_path, port=29633)
    for r in res:
        assert r[0] == 104.25 and int(r[1]) == 2
    allnan = np.full(5, np.nan)
    res = _run(allnan, tmp_path, port=29635)
    for r in res:
        assert np.isnan(r[0]) and int(r[1]) == -1
