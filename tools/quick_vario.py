"""Ad-hoc GPU timing of the variogram kernel (development aid, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stif_b200 import _native

ctx = _native.Context(0)
for it in (2048, 8192):
    rate, ms = ctx.bench_dfma(it)
    print(f"dfma peak: {rate:.4e} lane-instr/s  ({ms:.3f} ms)  = {2*rate/1e12:.2f} TFLOP/s")
for n in (100_000, 300_000):
    rng = np.random.default_rng(20261019)
    x = rng.uniform(0, 100, n); y = rng.uniform(0, 100, n)
    t = rng.integers(0, 365, n).astype(np.float64); v = rng.normal(size=n)
    for flags in (1, 3):
        for rep in range(3):
            t0 = time.time()
            c, s = ctx.variogram_host(x, y, t, v, 30.0, 30, 30, 30, flags=flags)
            wall = time.time() - t0
            ph = ctx.last_phase_ms()
        st = ctx.variogram_stats()
        pairs = n * (n - 1) // 2
        print(f"n={n} flags={flags} wall={wall*1e3:.2f} ms phases(ms)={['%.3f'%p for p in ph]} "
              f"pairs/s(kernel)={pairs/(ph[1]*1e-3):.3e} evaluated={st['pairs_evaluated']/pairs:.3f} "
              f"in_window={st['pairs_in_window']/pairs:.4f} binned={c.sum()/pairs:.4f}")
