"""Small kriging + variogram run for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stif_b200 import _native
ctx = _native.Context(0)
rng = np.random.default_rng(3)
n = 6000
x = rng.uniform(0, 100, n); y = rng.uniform(0, 100, n); t = rng.uniform(0, 365, n)
z = rng.normal(size=n)
c, s = ctx.variogram_host(x, y, np.floor(t), z, 30.0, 30.0, 30, 30)
print("variogram", int(c.sum()))
for flags in (_native.VARIO_PEEL | _native.VARIO_DETERMINISTIC, _native.VARIO_PEEL | _native.VARIO_SPACE_SORT,
              _native.VARIO_PEEL | _native.VARIO_NO_CULL | _native.VARIO_DETERMINISTIC):
    c2, s2 = ctx.variogram_host(x, y, np.floor(t), z, 30.0, 30.0, 30, 30, flags=flags)
    assert np.array_equal(c, c2)
c3, s3 = ctx.variogram_host(x, y, np.floor(t), z + 2.0, 30.0, 30.0, 12, 9, mode=1, mean=float(np.mean(z + 2.0)),
                            flags=_native.VARIO_PEEL | _native.VARIO_DETERMINISTIC)
print("covariogram (deterministic)", int(c3.sum()))
cs, ss = ctx.variogram_sampled_host(x, y, np.floor(t), z, 30.0, 30.0, 30, 30, 200000, seed=1)
print("sampled", int(cs.sum()))
tx = rng.uniform(0, 100, 300); ty = rng.uniform(0, 100, 300); tt = rng.uniform(30, 365, 300)
ctx.kriging_set_obs(x, y, t, z)
cb, sb = ctx.variogram_bound(np.arange(100, 4000), 30.0, 30.0, 30, 30)
print("variogram of a bound subset", int(cb.sum()))
pairs = ctx.kriging_neighbour_pairs(tx[:50], ty[:50], tt[:50], 20.0, 20.0)
print("neighbour pairs", len(pairs))
for models, params, k in [(("sum_metric","spherical","spherical","spherical"), [400,60,500,0.4,0.2,0.3,0.03,0.03,0.03,5.0], 50),
                          (("sum_metric","spherical","spherical","spherical"), [40,20,50,0.4,0.2,0.3,0.03,0.03,0.03,5.0], 70),
                          (("sum_metric","spherical","spherical","spherical"), [400,60,500,0.4,0.2,0.3,-10.0,-5.0,15.09,5.0], 21),
                          (("product_sum","spherical","spherical","spherical"), [300,60,0.8,0.7,0.15,0.10,0.4], 50),
                          (("product_sum","spherical","gaussian","spherical"), [300,60,0.8,0.7,0.15,0.10,0.4], 100),
                          (("sum","spherical","gaussian","spherical"), [300,60,0.8,0.7,0.15,0.10], 9)]:
    ctx.kriging_set_model(*models, params)
    m, sd, w, g, idx, nu = ctx.kriging_predict_host(tx, ty, tt, 1, k, 60.0, 60.0, want_weights=True)
    print(models[0], k, float(np.nansum(m)), ctx.kriging_solver_stats())
# duplicate observations: rank-deficient systems -> minimum-norm kernel
x2, y2, t2 = x.copy(), y.copy(), t.copy()
x2[3000:] = x2[:3000]; y2[3000:] = y2[:3000]; t2[3000:] = t2[:3000]
ctx.kriging_set_obs(x2, y2, t2, z)
ctx.kriging_set_model("sum_metric", "spherical", "spherical", "spherical", [400,60,500,0.4,0.2,0.3,0.03,0.03,0.03,5.0])
m, sd = ctx.kriging_predict_host(tx, ty, tt, 1, 24, 60.0, 60.0)
print("duplicates", float(np.nansum(m)), ctx.kriging_stats())
ctx.close()
