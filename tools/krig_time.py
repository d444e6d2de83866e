"""Ad-hoc GPU timing of the kriging path on the bench configurations (development aid).
CFG=c4|c5 NT=<targets> REPS=<n> FLAGS=<solver flags: 2 LU, 4 try Cholesky, 8 try LDL'> STIF_LIB=<other build>"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from stif_b200 import _native
if os.environ.get("STIF_LIB"):
    _native.LIB_PATH = os.path.abspath(os.environ["STIF_LIB"])
ctx = _native.Context(0)
for key in os.environ.get("CFG", "c4,c5").split(","):
    cfg = bench.KRIG[key]
    n_t = int(os.environ.get("NT", cfg["targets_per_gpu"]))
    (ox, oy, ot, oz), (tx, ty, tt) = bench.synth_kriging(cfg["n_obs"], n_t, 0, cfg["seed"], cfg["indicator"])
    ctx.kriging_set_obs(ox, oy, ot, oz)
    ctx.kriging_set_model(*cfg["model"], cfg["params"])
    best = None
    for rep in range(int(os.environ.get("REPS", 3))):
        mean, std = ctx.kriging_predict_host(tx, ty, tt, 1, cfg["k"], cfg["radii"][0], cfg["radii"][1],
                                             flags=_native.KRIG_F32_STORAGE | int(os.environ.get("FLAGS", 0)))
        ph = ctx.last_phase_ms()
        if best is None or ph[1] + ph[2] < best[1] + best[2]:
            best = ph
    st = ctx.kriging_stats()
    print(f"{key} k={cfg['k']} targets={n_t} search={best[1]:.3f} solve={best[2]:.3f} ms -> {n_t/((best[1]+best[2])*1e-3):.4e} pred/s"
          f" examined/target={st['candidates_examined']/n_t:.1f} in_window={st['candidates_in_window']/n_t:.1f}"
          f" to_lu={ctx.kriging_solver_stats()['first_pass_to_lu']}"
          f" checksum={float(np.nansum(mean)):.9e} nan={int(np.isnan(mean).sum())}", flush=True)
