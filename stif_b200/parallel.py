"""Multi-GPU layer: one process per GPU, torch.distributed for the plumbing.

Variogram: the tile-pair work list is partitioned over ranks inside the kernel
(`part`/`nparts`); the observations are uploaded 1/N per rank and all-gathered over NVLink,
and the only exchange of results is ONE all-reduce of a fused float64 buffer holding the
(n_space_bins x n_time_bins) count and sum histograms (14.4 KB at 30x30 -> latency bound;
NCCL over NVLink on GPUs, gloo on CPU in the tests).  Kriging: targets are sharded contiguously, observations are
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
    """Sum uint64 counts and float64 sums over all ranks with ONE all-reduce of one float64
    buffer [counts | sums].  Counts travel as float64 -- the dtype of the reference's own `norm`
    array -- which is exact below 2^53 (n(n-1)/2 < 2^53 for any n < 1.3e8), so they stay
    bit-exact."""
    rank, ws = world()
    if ws == 1:
        return counts, sums
    import torch
    import torch.distributed as dist
    dev = _device_for_collectives()
    shape = counts.shape
    c = np.ascontiguousarray(counts).reshape(-1)
    if int(c.max(initial=0)) >= (1 << 53) // ws:
        raise OverflowError("pair counts beyond 2^53 cannot be all-reduced exactly as float64")
    buf = np.concatenate([c.astype(np.float64), np.ascontiguousarray(sums, dtype=np.float64).reshape(-1)])
    tbuf = torch.from_numpy(buf).to(dev)
    dist.all_reduce(tbuf, op=dist.ReduceOp.SUM)
    out = tbuf.cpu().numpy()
    nb = c.size
    return out[:nb].astype(np.uint64).reshape(shape), out[nb:].reshape(shape)


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
    """All-pairs histograms, work split over the ranks of the default process group.

    NCCL (one GPU per rank): every rank uploads only ITS 1/N of the observations (the host arrays
    are replicated, the PCIe traffic is not), the shards meet in an all-gather over NVLink, each
    rank runs the pair kernel on its tile partition of the device-resident set, and counts and
    sums of all ranks meet in one fused float64 all-reduce; 16 B per bin travel back to the host.
    Other backends (gloo in the CPU tests): the host entry point per rank + the same all-reduce."""
    from . import _native
    rank, ws = world()
    if flags is None:
        flags = _native.VARIO_PEEL
    if ws == 1:
        return ctx.variogram_host(x, y, t, v, smax, tmax, ns, nt, mode=mode, mean=mean, flags=flags)
    import torch
    import torch.distributed as dist
    if "nccl" not in str(dist.get_backend()):
        counts, sums = ctx.variogram_host(x, y, t, v, smax, tmax, ns, nt, mode=mode, mean=mean,
                                          flags=flags, part=rank, nparts=ws)
        return allreduce_histograms(counts, sums)
    dev = torch.device("cuda", torch.cuda.current_device())
    n = len(x)
    nb = int(ns) * int(nt)
    shard = (n + ws - 1) // ws                      # equal shard length: rank r owns [r*shard, (r+1)*shard)
    b, e = min(rank * shard, n), min((rank + 1) * shard, n)
    full = torch.empty((4, ws * shard), dtype=torch.float64, device=dev)
    mine = torch.zeros((4, shard), dtype=torch.float64, device=dev)
    for i, a in enumerate((x, y, t, v)):
        if e > b:                                   # async DMA when the caller's arrays are pinned
            mine[i, : e - b].copy_(torch.from_numpy(_native._f64(a)[b:e]), non_blocking=True)
    for i in range(4):                              # NVLink: 8 B x n per field, every rank ends with all of it
        dist.all_gather_into_tensor(full[i], mine[i])
    fused = torch.empty(2 * nb, dtype=torch.float64, device=dev)
    ctx.variogram_dev(full[0].data_ptr(), full[1].data_ptr(), full[2].data_ptr(), full[3].data_ptr(), n,
                      smax, tmax, ns, nt, fused.data_ptr(), fused.data_ptr() + 8 * nb, mode=mode, mean=mean,
                      flags=flags | _native.VARIO_COUNTS_F64, part=rank, nparts=ws,
                      stream=torch.cuda.current_stream().cuda_stream)
    dist.all_reduce(fused)
    out = fused.cpu().numpy()
    shape = (int(ns), int(nt))
    return out[:nb].astype(np.uint64).reshape(shape), out[nb:].reshape(shape)


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
