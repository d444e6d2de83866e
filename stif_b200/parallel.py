"""Multi-GPU layer: one process per GPU, torch.distributed for the plumbing.

Variogram: the tile-pair work list is partitioned over ranks inside the kernel
(`part`/`nparts`); the only exchange is one all-reduce of the (n_space_bins x n_time_bins)
count and sum histograms (14.4 KB at 30x30 -> latency bound; NCCL over NVLink on GPUs,
gloo on CPU in the tests).  Kriging: targets are sharded contiguously, observations are
replicated, results are gathered -- no data-path collective.

Nothing here is needed for single-GPU use; with torch.distributed uninitialised every
helper degrades to world_size 1.
"""
import numpy as np


def world():
    """(rank, world_size) of the default process group, (0, 1) if there is none."""
    try:
        import torch.distributed as dist
    except Exception:  # torch absent: single process
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous, balanced [begin, end) of `n` items for `rank` (first n % world get +1)."""
    base, extra = divmod(int(n), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _device_for_collectives():
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    if "nccl" in str(backend):
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_histograms(counts, sums):
    """Sum uint64 counts and float64 sums over all ranks (bit-exact for the counts)."""
    rank, ws = world()
    if ws == 1:
        return counts, sums
    import torch
    import torch.distributed as dist
    dev = _device_for_collectives()
    shape = counts.shape
    c = torch.from_numpy(np.ascontiguousarray(counts).view(np.int64).reshape(-1).copy()).to(dev)
    s = torch.from_numpy(np.ascontiguousarray(sums, dtype=np.float64).reshape(-1).copy()).to(dev)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return (c.cpu().numpy().view(np.uint64).reshape(shape), s.cpu().numpy().reshape(shape))


def gather_concat(local, n_total):
    """Concatenate per-rank 1-D/2-D arrays (contiguous shards in rank order) on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    import torch
    import torch.distributed as dist
    dev = _device_for_collectives()
    local = np.ascontiguousarray(local)
    tail = local.shape[1:]
    sizes = [shard_bounds(n_total, r, ws) for r in range(ws)]
    maxlen = max(e - b for b, e in sizes)
    pad = np.zeros((maxlen,) + tail, dtype=local.dtype)
    pad[: len(local)] = local
    view_dtype = {4: np.int32, 8: np.int64}[local.dtype.itemsize] if local.dtype.kind in "u" else local.dtype
    t = torch.from_numpy(pad.view(view_dtype)).to(dev)
    outs = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(outs, t)
    parts = [o.cpu().numpy().view(local.dtype)[: e - b] for o, (b, e) in zip(outs, sizes)]
    return np.concatenate(parts, axis=0)


def variogram_all_ranks(ctx, x, y, t, v, smax, tmax, ns, nt, mode, mean, flags=None):
    """All-pairs histograms, work split over the ranks of the default process group."""
    from . import _native
    rank, ws = world()
    if flags is None:
        flags = _native.VARIO_PEEL
    counts, sums = ctx.variogram_host(x, y, t, v, smax, tmax, ns, nt, mode=mode, mean=mean,
                                      flags=flags, part=rank, nparts=ws)
    return allreduce_histograms(counts, sums)


def variogram_sampled_all_ranks(ctx, x, y, t, v, smax, tmax, ns, nt, mode, mean, n_samples):
    """Sampled-pair branch (el_max): every rank draws n_samples/world pairs."""
    rank, ws = world()
    b, e = shard_bounds(n_samples, rank, ws)
    counts, sums = ctx.variogram_sampled_host(x, y, t, v, smax, tmax, ns, nt, e - b, mode=mode,
                                              mean=mean, seed_offset=rank)
    return allreduce_histograms(counts, sums)


def kriging_all_ranks(predict_fn, n_targets):
    """Run predict_fn(begin, end) -> tuple of arrays on this rank's target shard and gather."""
    rank, ws = world()
    if ws == 1:
        return predict_fn(0, n_targets)
    b, e = shard_bounds(n_targets, rank, ws)
    local = predict_fn(b, e)
    return tuple(gather_concat(a, n_targets) for a in local)
