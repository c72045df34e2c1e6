"""Candidate sharding across GPUs (one process per GPU, torch.distributed).

Candidates are independent, so the path shards with no data-path collective: rank r evaluates the contiguous
slice ``shard_range(B, r, world)`` of the batch on its own GPU.  The only exchange is the final reduction the
callers want - the best lap and which candidate produced it - done as a local argmin followed by one all-gather of
a (lap, global index) pair per rank (16 bytes each; NCCL over NVLink on GPUs, gloo on CPU in the tests) and an
argmin over the gathered pairs on every rank.  `all_gather_laps` is the full-vector variant (BASELINE config 3).
"""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """[lo, hi) of the candidates rank `rank` owns: contiguous, balanced to within one candidate."""
    lo = (B * rank) // world
    hi = (B * (rank + 1)) // world
    return lo, hi


def global_argmin(local_best_lap, local_best_idx, shard_lo, group=None):
    """local_best_lap: 0-d/1-element float64 tensor (NaN if the shard had no valid candidate), local_best_idx:
    index within the shard (-1 if none).  Returns (best_lap, best_global_idx) tensors, identical on every rank.
    Ties resolve to the lowest global index, as a single-GPU argmin over the whole batch would."""
    dev = local_best_lap.device
    lap = local_best_lap.reshape(1).to(torch.float64)
    idx = local_best_idx.reshape(1).to(torch.int64)
    gidx = torch.where(idx >= 0, idx + int(shard_lo), torch.full_like(idx, -1))
    pair = torch.stack([lap, gidx.to(torch.float64)]).reshape(2)   # indices < 2^53 are exact in float64
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return lap[0], gidx[0]
    world = dist.get_world_size(group)
    gathered = torch.empty(2 * world, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gathered, pair, group=group)
    laps, idxs = gathered[0::2], gathered[1::2].to(torch.int64)
    valid = (idxs >= 0) & ~torch.isnan(laps)
    key = torch.where(valid, laps, torch.full_like(laps, float("inf")))
    best = torch.min(key)
    cand = torch.where(valid & (key == best), idxs, torch.full_like(idxs, torch.iinfo(torch.int64).max))
    best_idx = torch.min(cand)
    none = ~valid.any()
    return (torch.where(none, torch.full_like(best, float("nan")), best),
            torch.where(none, torch.full_like(best_idx, -1), best_idx))


def all_gather_laps(local_laps, B, group=None):
    """Full lap vector on every rank from per-rank slices (ragged shards allowed): pads to the largest shard,
    all-gathers, and re-assembles in candidate order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_laps
    world = dist.get_world_size(group)
    sizes = [shard_range(B, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    buf = torch.full((width,), float("nan"), dtype=local_laps.dtype, device=local_laps.device)
    buf[: local_laps.numel()] = local_laps
    out = torch.empty(world * width, dtype=local_laps.dtype, device=local_laps.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


class DeviceArgmin:
    """The same reduction on the GPU with three small kernels of the library and ONE collective of 16 bytes per rank
    (`global_argmin` above spends ~10 eager torch launches on it): local argmin over the shard -> {lap, global index}
    pair (sto_argmin_pair_f64) -> ncclAllGather of the pairs -> argmin over the pairs (sto_argmin_pairs_f64).  Buffers
    are allocated once; nothing synchronises.  Identical on every rank, ties to the lowest global index."""

    def __init__(self, device, group=None):
        import ctypes as C
        from . import _lib
        self._C, self._lib_mod, self.lib = C, _lib, _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.pair = torch.empty(2, dtype=torch.float64, device=device)
        self.scratch = torch.empty(2, dtype=torch.float64, device=device)
        self.gathered = torch.empty(2 * self.world, dtype=torch.float64, device=device)
        self.best = torch.empty(1, dtype=torch.float64, device=device)
        self.idx = torch.empty(1, dtype=torch.int64, device=device)

    def __call__(self, lap, status, shard_lo):
        """lap[B], status[B] (or None): this rank's shard (device).  Returns (best_lap[1], best_global_idx[1])."""
        C, chk = self._C, self._lib_mod.check
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        chk(self.lib.sto_argmin_pair_f64(p(lap), p(status), lap.numel(), int(shard_lo), p(self.pair), p(self.scratch), st))
        src = self.pair
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.pair, group=self.group)
            src = self.gathered
        chk(self.lib.sto_argmin_pairs_f64(p(src), self.world, p(self.best), p(self.idx), st))
        return self.best, self.idx
