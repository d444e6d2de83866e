"""Whole-workload golden histograms for C2 (100k observations) and C3 (1M observations).

TEST INFRASTRUCTURE.  Runs the C restatement of the reference's pair loop
(oracle_variogram.c, pinned bit-exactly against the unmodified numba reference by
tests/test_oracle_cpu.py) over EVERY pair of the bench workloads, row-block-parallel over the
host cores, and commits the 30 x 30 counts and sums as tests/golden/variogram_c2.npz /
variogram_c3.npz.  tests/test_fullsize_gpu.py and bench.py (at every N) compare the GPU
result with these files: counts bit-exact, sums to 1e-12 relative.

Counts are integers and independent of the block split.  The sums are accumulated
sequentially inside a row block (the reference's order) and the block partials are added in
block order with math.fsum-free plain float64 adds; that differs from the reference's single
sequential chain by rounding only (relative 1e-13 at these sizes), which the 1e-12 tolerance
of the north star covers.

    python -m oracle.gen_fullsize c2        # ~30 core-seconds
    python -m oracle.gen_fullsize c3        # ~50 core-minutes (8 cores: ~7 min)
"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import variogram as ov   # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SIZES = {"c2": 100_000, "c3": 1_000_000}
VARIO = dict(space_dist_max=30.0, time_dist_max=30.0, n_space_bins=30, n_time_bins=30)


def synth_variogram(n):
    """SURVEY 8(d) C2/C3 generator (the same as bench.py's)."""
    rng = np.random.default_rng(20261019)
    x = rng.uniform(0, 100, n)
    y = rng.uniform(0, 100, n)
    t = rng.integers(0, 365, n).astype(np.float64)
    v = rng.normal(size=n)
    return x, y, t, v


def row_blocks(n, n_blocks):
    """Row ranges [b, e) of the upper triangle with about equal pair counts."""
    total = n * (n - 1) / 2
    edges = [0]
    for k in range(1, n_blocks):
        # rows 0..r-1 hold r*n - r(r+1)/2 pairs; solve for the k/n_blocks quantile
        target = total * k / n_blocks
        r = int(n - 0.5 - np.sqrt((n - 0.5) ** 2 - 2 * target))
        edges.append(max(edges[-1], min(n, r)))
    edges.append(n)
    return [(b, e) for b, e in zip(edges[:-1], edges[1:]) if e > b]


def run(name, threads=None):
    n = SIZES[name]
    x, y, t, v = synth_variogram(n)
    space = np.column_stack([x, y])
    threads = threads or os.cpu_count() or 1
    blocks = row_blocks(n, 64 if n <= 200_000 else 512)
    ov.raw(space[:100], t[:100], v[:100], 30.0, 30.0, 30, 30)   # build + load once

    def work(be):
        b, e = be
        # the PEEL flag only affects pair (0, 1), which lives in the first block
        return ov.raw(space, t, v, VARIO["space_dist_max"], VARIO["time_dist_max"], 30, 30,
                      i_begin=b, i_end=e)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:      # ctypes releases the GIL during the C loop
        parts = list(ex.map(work, blocks))
    dt = time.perf_counter() - t0
    sums = np.zeros((30, 30))
    counts = np.zeros((30, 30), dtype=np.int64)
    for h, c in parts:
        sums += h
        counts += c.astype(np.int64)
    pairs = n * (n - 1) // 2
    out = os.path.join(GOLDEN, f"variogram_{name}.npz")
    np.savez_compressed(out, counts=counts, sums=sums, n=np.int64(n), pairs=np.int64(pairs),
                        seed=np.int64(20261019), smax=30.0, tmax=30.0,
                        row_blocks=np.array(blocks, dtype=np.int64))
    print(f"{name}: n={n} pairs={pairs} binned={int(counts.sum())} in {dt:.1f} s on {threads} threads "
          f"({pairs / dt / threads:.3e} pairs/s/thread) -> {out}")


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["c2"]):
        run(name)
