"""Timing of the verbatim-gamma Cholesky variant (cancelling marginal nuggets, as unconstrained fits produce)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stif_b200 import _native
ctx = _native.Context(0)
n_obs = 1_000_000; n_t = 1_250_000
rng = np.random.default_rng(20261020)
x = rng.uniform(0, 1000, n_obs); y = rng.uniform(0, 1000, n_obs); t = rng.uniform(0, 365, n_obs)
z = np.sin(x/150)+np.cos(y/200)+0.5*np.sin(t/30)+0.3*rng.normal(size=n_obs)
tx = rng.uniform(0, 1000, n_t); ty = rng.uniform(0, 1000, n_t); tt = rng.uniform(30, 365, n_t)
ctx.kriging_set_obs(x, y, t, z)
for name, params in [("plain", [400,60,500,0.4,0.2,0.3,0.03,0.03,0.03,5.0]),
                     ("cancelling nuggets", [400,60,500,0.4,0.2,0.3,-1000.0,-500.0,1500.09,5.0]),
                     ("gaussian metric", [400,60,500,0.4,0.2,0.3,0.03,0.03,0.03,5.0])]:
    models = ("sum_metric","spherical","spherical","gaussian" if name.startswith("gauss") else "spherical")
    ctx.kriging_set_model(*models, params)
    for rep in range(2):
        mean, std = ctx.kriging_predict_host(tx, ty, tt, 1, 50, 30.0, 30.0)
        ph = ctx.last_phase_ms()
    print(f"{name}: search={ph[1]:.2f} solve={ph[2]:.2f} ms", ctx.kriging_solver_stats(), "nan", np.isnan(mean).sum())
