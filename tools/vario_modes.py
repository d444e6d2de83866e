"""Ad-hoc GPU timing: time-sorted vs space-time-sorted work lists (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stif_b200 import _native
if os.environ.get("STIF_LIB"):
    _native.LIB_PATH = os.path.abspath(os.environ["STIF_LIB"])
ctx = _native.Context(0)
for n in [int(a) for a in os.environ.get("NS", "100000,300000,1000000").split(",")]:
    rng = np.random.default_rng(20261019)
    x = rng.uniform(0, 100, n); y = rng.uniform(0, 100, n)
    t = rng.integers(0, 365, n).astype(np.float64); v = rng.normal(size=n)
    ref = None
    for name, fl in (("time", _native.VARIO_PEEL | _native.VARIO_TIME_SORT), ("space", _native.VARIO_PEEL | _native.VARIO_SPACE_SORT)):
        for rep in range(3):
            c, s = ctx.variogram_host(x, y, t, v, 30.0, 30, 30, 30, flags=fl, part=0, nparts=int(os.environ.get("NPARTS", 1)))
            ph = ctx.last_phase_ms()
        st = ctx.variogram_stats()
        pairs = n * (n - 1) // 2
        if ref is None: ref = c
        print(f"n={n} {name:5s} phases(ms)={['%.3f' % p for p in ph]} evaluated={st['pairs_evaluated']/pairs:.4f} same_counts={np.array_equal(ref, c)}")
