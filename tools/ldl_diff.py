"""Compare two builds of the library target by target on the C5 configuration (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from stif_b200 import _native
cfg = bench.KRIG["c5"]
n_t = int(os.environ.get("NT", 40000))
(ox, oy, ot, oz), (tx, ty, tt) = bench.synth_kriging(cfg["n_obs"], n_t, 0, cfg["seed"], cfg["indicator"])
out = {}
for name, lib in (("new", None), ("old", os.environ["STIF_LIB_OLD"])):
    if lib:
        _native.LIB_PATH = os.path.abspath(lib)
        _native._lib = None if hasattr(_native, "_lib") else None
    ctx = _native.Context(0)
    ctx.kriging_set_obs(ox, oy, ot, oz)
    ctx.kriging_set_model(*cfg["model"], cfg["params"])
    res = []
    for rep in range(2):
        m, s, w, g, idx, nu = ctx.kriging_predict_host(tx, ty, tt, 1, cfg["k"], cfg["radii"][0], cfg["radii"][1], flags=0,
                                                        want_weights=True)
        res.append((m.copy(), w.copy()))
    print(name, "run-to-run identical:", np.array_equal(res[0][0], res[1][0], equal_nan=True), flush=True)
    out[name] = res[0]
    del ctx
m1, w1 = out["new"]; m2, w2 = out["old"]
diff = np.abs(m1 - m2)
sc = np.sum(np.abs(w2.astype(np.float64)) * 1.0, axis=1) + 1e-300
bad = np.nonzero(diff > 0)[0]
print("targets differing:", len(bad), "of", n_t, " max |dmean|/sum|w|:", float((diff / sc).max()))
print("weights: max |dw|/rowmax:", float((np.abs(w1 - w2) / np.abs(w2).max(axis=1, keepdims=True)).max()))
