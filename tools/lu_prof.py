import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from stif_b200 import _native
ctx = _native.Context(0)
n_obs = 1_000_000; n_t = 60000
rng = np.random.default_rng(20261020)
x = rng.uniform(0, 1000, n_obs); y = rng.uniform(0, 1000, n_obs); t = rng.uniform(0, 365, n_obs)
z = rng.normal(size=n_obs)
tx = rng.uniform(0, 1000, n_t); ty = rng.uniform(0, 1000, n_t); tt = rng.uniform(30, 365, n_t)
ctx.kriging_set_obs(x, y, t, z)
ctx.kriging_set_model("product_sum","spherical","spherical","spherical", [300,60,0.8,0.7,0.15,0.10,0.4])
for k in (50, 100):
    mean, std = ctx.kriging_predict_host(tx[:n_t if k==50 else 15000], ty[:n_t if k==50 else 15000], tt[:n_t if k==50 else 15000], 1, k, 30.0, 30.0)
    print(k, ctx.last_phase_ms())
