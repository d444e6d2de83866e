"""Ad-hoc GPU timing of the variogram path for A/B runs (development aid, not the bench).
STIF_LIB=path/to/other.so selects another build of the library."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stif_b200 import _native
if os.environ.get("STIF_LIB"):
    _native.LIB_PATH = os.path.abspath(os.environ["STIF_LIB"])
ctx = _native.Context(0)
sizes = [int(s) for s in os.environ.get("SIZES", "100000,300000,1000000").split(",")]
for n in sizes:
    rng = np.random.default_rng(20261019)
    x = rng.uniform(0, 100, n); y = rng.uniform(0, 100, n)
    t = rng.integers(0, 365, n).astype(np.float64); v = rng.normal(size=n)
    ker, prep, hashes = [], [], set()
    for rep in range(int(os.environ.get("REPS", 5))):
        c, s = ctx.variogram_host(x, y, t, v, 30.0, 30, 30, 30, flags=int(os.environ.get("FLAGS", 1)))
        ph = ctx.last_phase_ms()
        prep.append(ph[0]); ker.append(ph[1])
        hashes.add(hashlib.sha256(s.tobytes()).hexdigest()[:12])
    st = ctx.variogram_stats()
    pairs = n * (n - 1) // 2
    g = None
    gp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                      {100000: "variogram_c2.npz", 1000000: "variogram_c3.npz"}.get(n, "none"))
    ok = ""
    if os.path.exists(gp):
        g = np.load(gp)
        ok = f" counts_ok={np.array_equal(c.astype(np.int64), g['counts'])} sums_rel={np.max(np.abs(s - g['sums']) / np.abs(g['sums'])):.2e}"
    print(f"n={n} kernel min/med {min(ker):.3f}/{sorted(ker)[len(ker)//2]:.3f} ms prepare {min(prep):.3f} ms "
          f"evaluated={st['pairs_evaluated']/pairs:.4f} items={st['items']} distinct_sum_hashes={len(hashes)}{ok}", flush=True)
