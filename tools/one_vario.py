import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stif_b200 import _native
ctx = _native.Context(0)
n = int(os.environ.get("N", 100000)); flags = int(os.environ.get("FLAGS", 1)); nparts = int(os.environ.get("NPARTS", 1))
rng = np.random.default_rng(20261019)
x = rng.uniform(0, 100, n); y = rng.uniform(0, 100, n)
t = rng.integers(0, 365, n).astype(np.float64); v = rng.normal(size=n)
for rep in range(2):
    c, s = ctx.variogram_host(x, y, t, v, 30.0, 30, 30, 30, flags=flags, part=0, nparts=nparts)
print(ctx.last_phase_ms(), c.sum())
