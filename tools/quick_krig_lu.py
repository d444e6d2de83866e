"""Ad-hoc GPU timing of the pivoted-LU solve path (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stif_b200 import _native
ctx = _native.Context(0)
n_obs = 1_000_000; n_t = int(os.environ.get("NT", 400_000))
rng = np.random.default_rng(20261020)
x = rng.uniform(0, 1000, n_obs); y = rng.uniform(0, 1000, n_obs); t = rng.uniform(0, 365, n_obs)
z = np.sin(x/150)+np.cos(y/200)+0.5*np.sin(t/30)+0.3*rng.normal(size=n_obs)
tx = rng.uniform(0, 1000, n_t); ty = rng.uniform(0, 1000, n_t); tt = rng.uniform(30, 365, n_t)
ctx.kriging_set_obs(x, y, t, z)
for models, params, k, flags in [
        (("sum_metric","spherical","spherical","spherical"), [400,60,500,0.4,0.2,0.3,0.03,0.03,0.03,5.0], 50, _native.KRIG_F32_STORAGE | _native.KRIG_FORCE_LU),
        (("product_sum","spherical","spherical","spherical"), [300,60,0.8,0.7,0.15,0.10,0.4], 50, _native.KRIG_F32_STORAGE),
        (("product_sum","spherical","spherical","spherical"), [300,60,0.8,0.7,0.15,0.10,0.4], 100, _native.KRIG_F32_STORAGE),
        (("product_sum","spherical","spherical","spherical"), [300,60,0.8,0.7,0.15,0.10,0.4], 20, _native.KRIG_F32_STORAGE)]:
    ctx.kriging_set_model(*models, params)
    nt = n_t if k <= 50 else n_t // 4
    for rep in range(2):
        mean, std = ctx.kriging_predict_host(tx[:nt], ty[:nt], tt[:nt], 1, k, 30.0, 30.0, flags=flags)
        ph = ctx.last_phase_ms()
    st = ctx.kriging_stats()
    print(f"{models[0]} k={k} LU targets={nt} search={ph[1]:.2f} solve={ph[2]:.2f} ms  -> {nt/(ph[2]*1e-3):.3e} solves/s singular={st['targets_singular']} nan={np.isnan(mean).sum()}")
