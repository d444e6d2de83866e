"""Wall-clock of the Predictor API on the PM10-sized workload (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from stif_b200 import Data, Predictor

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "pm10.npz"))
df = pd.DataFrame({k: g[k] for k in g.files})
print(df.shape, list(df.columns))
data = Data(df, space_cols=["x", "y"], time_col="time", predictand_col="pm10")
p = Predictor(data)
for rep in range(3):
    t0 = time.perf_counter(); p.calc_empirical_variogram(space_dist_max=6e5, time_dist_max=7, n_space_bins=10, n_time_bins=7); t1 = time.perf_counter()
    p.fit_variogram_model(st_model="product_sum"); t2 = time.perf_counter()
    sp, tm = data.space_coords, data.time_coords
    mean, std = p.get_kriging_prediction(sp, tm, max_kriging_points=50, space_dist_max=2e5, leave_out_idxs=np.arange(len(tm))); t3 = time.perf_counter()
    p.fit_variogram_model(st_model="sum_metric"); t4 = time.perf_counter()
    mean2, std2 = p.get_kriging_prediction(sp, tm, max_kriging_points=50, space_dist_max=2e5, leave_out_idxs=np.arange(len(tm))); t5 = time.perf_counter()
    print(f"variogram {1e3*(t1-t0):.1f} ms | fit product_sum {1e3*(t2-t1):.0f} ms | kriging {len(tm)} targets (product_sum) {1e3*(t3-t2):.1f} ms"
          f" | fit sum_metric {1e3*(t4-t3):.0f} ms | kriging (sum_metric) {1e3*(t5-t4):.1f} ms | nan {np.isnan(mean).sum()} {np.isnan(mean2).sum()}")
    print("   stats", p.context.kriging_stats(), p.context.kriging_solver_stats(), p._variogram_fit.x)
