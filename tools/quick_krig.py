"""Ad-hoc GPU timing of the kriging path (development aid, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stif_b200 import _native
if os.environ.get("STIF_LIB"):            # A/B against another build of the library
    _native.LIB_PATH = os.path.abspath(os.environ["STIF_LIB"])
ctx = _native.Context(0)
n_obs = int(os.environ.get("NOBS", 1_000_000)); n_t = int(os.environ.get("NT", 1_000_000))
rng = np.random.default_rng(20261020)
x = rng.uniform(0, 1000, n_obs); y = rng.uniform(0, 1000, n_obs); t = rng.uniform(0, 365, n_obs)
z = np.sin(x/150)+np.cos(y/200)+0.5*np.sin(t/30)+0.3*rng.normal(size=n_obs)
tx = rng.uniform(0, 1000, n_t); ty = rng.uniform(0, 1000, n_t); tt = rng.uniform(30, 365, n_t)
ctx.kriging_set_obs(x, y, t, z)
for models, params, k in [(("sum_metric","spherical","spherical","spherical"), [400,60,500,0.4,0.2,0.3,0.03,0.03,0.03,5.0], 50),
                          (("product_sum","spherical","spherical","spherical"), [300,60,0.8,0.7,0.15,0.10,0.4], 100)]:
    ctx.kriging_set_model(*models, params)
    nt = n_t if k == 50 else n_t // 4
    if k != 50 and os.environ.get("ONLY50"): continue
    for rep in range(int(os.environ.get("REPS", 2))):
        t0 = time.time()
        mean, std = ctx.kriging_predict_host(tx[:nt], ty[:nt], tt[:nt], 1, k, 30.0, 30.0)
        wall = time.time() - t0
        ph = ctx.last_phase_ms()
    st = ctx.kriging_stats()
    print(f"{models[0]} k={k} targets={nt} wall={wall*1e3:.1f} ms phases(ms) grid={ph[0]:.2f} search={ph[1]:.2f} solve={ph[2]:.2f}"
          f" pred/s(dev)={nt/((ph[1]+ph[2])*1e-3):.3e} cand/target={st['candidates_in_window']/nt:.1f}"
          f" examined/target={st['candidates_examined']/nt:.1f} singular={st['targets_singular']} nan={np.isnan(mean).sum()}")
