"""Ad-hoc GPU timing of the solvers of the indefinite (product_sum) systems: the pivoted LU against
the Bunch-Kaufman LDL' kernel at k = 20 / 50 / 100 (development aid).  STIF_LIB=<other build>."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stif_b200 import _native
if os.environ.get("STIF_LIB"):
    _native.LIB_PATH = os.path.abspath(os.environ["STIF_LIB"])
ctx = _native.Context(0)
n_obs = 1_000_000; n_t = int(os.environ.get("NT", 400_000))
rng = np.random.default_rng(20261020)
x = rng.uniform(0, 1000, n_obs); y = rng.uniform(0, 1000, n_obs); t = rng.uniform(0, 365, n_obs)
z = np.sin(x/150)+np.cos(y/200)+0.5*np.sin(t/30)+0.3*rng.normal(size=n_obs)
tx = rng.uniform(0, 1000, n_t); ty = rng.uniform(0, 1000, n_t); tt = rng.uniform(30, 365, n_t)
ctx.kriging_set_obs(x, y, t, z)
PS = (("product_sum","spherical","spherical","spherical"), [300,60,0.8,0.7,0.15,0.10,0.4])
SM = (("sum_metric","spherical","spherical","spherical"), [400,60,500,0.4,0.2,0.3,0.03,0.03,0.03,5.0])
F32 = _native.KRIG_F32_STORAGE
for (models, params), k, flags, name in [
        (SM, 50, F32 | _native.KRIG_FORCE_LU, "LU"), (SM, 50, F32 | _native.KRIG_TRY_LDL, "LDL"), (SM, 50, F32, "CHOL"),
        (PS, 50, F32 | _native.KRIG_FORCE_LU, "LU"), (PS, 50, F32, "LDL"),
        (PS, 100, F32 | _native.KRIG_FORCE_LU, "LU"), (PS, 100, F32, "LDL"),
        (PS, 20, F32 | _native.KRIG_FORCE_LU, "LU"), (PS, 20, F32, "LDL")]:
    ctx.kriging_set_model(*models, params)
    nt = n_t if k <= 50 else n_t // 4
    for rep in range(2):
        mean, std = ctx.kriging_predict_host(tx[:nt], ty[:nt], tt[:nt], 1, k, 30.0, 30.0, flags=flags)
        ph = ctx.last_phase_ms()
    st, ss = ctx.kriging_stats(), ctx.kriging_solver_stats()
    print(f"{models[0]} k={k} {name} targets={nt} search={ph[1]:.2f} solve={ph[2]:.2f} ms  -> {nt/(ph[2]*1e-3):.3e} solves/s "
          f"singular={st['targets_singular']} to_lu={ss['first_pass_to_lu']} nan={np.isnan(mean).sum()} "
          f"checksum={float(np.nansum(mean)):.9e}", flush=True)
