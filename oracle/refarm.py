"""CPU baseline arm of bench.py -- TEST / BENCH INFRASTRUCTURE, never on the product path.

Times the reference's own implementation of the hot path on the host cores:

  kind "reference": the UNMODIFIED numba package (vendored by oracle/vendor_ref.py into
      oracle/_ref/, imported with a stub matplotlib by oracle/refload.py) through its public
      entry points -- `stif.utils.get_variogram(..., n_samples=None)` (serial by construction:
      1 core) and `Predictor.get_kriging_prediction` with an injected parameter vector (the
      pre-filter `Data.get_kriging_idxs` is single-threaded, `nd_kriging` is a numba `prange`
      over `numba.get_num_threads()` threads).  JIT time is measured by a warm-up call on a
      small input of the same dtypes and reported separately.
  kind "port": the C restatement (oracle/*.c), used only if the vendored package or numba
      cannot be imported on the box.
"""
import os
import time
import types

import numpy as np


def _load_reference():
    from . import refload
    if not refload.available(vendored_only=True):
        return None, "oracle/_ref/ is absent (run __graft_entry__.build() in the build container)"
    try:
        import numba  # noqa: F401
        return refload.load(vendored_only=True), None
    except Exception as exc:  # numba / llvmlite / pandas missing or incompatible
        return None, f"{type(exc).__name__}: {exc}"


def variogram_sample(x, y, t_int, v, n_sub, smax, tmax, ns, nt, repeats=1):
    """All-pairs variogram of the first n_sub observations, `repeats` times.  Returns a dict with
    pairs/s (rate is independent of n: 9.7e7 vs 9.0e7 pairs/s at 30k vs 100k)."""
    stif, why = _load_reference()
    space = np.column_stack([x[:n_sub], y[:n_sub]])
    tt = np.ascontiguousarray(t_int[:n_sub])
    vv = np.ascontiguousarray(v[:n_sub])
    pairs = n_sub * (n_sub - 1) // 2
    if stif is not None:
        from stif.utils import get_variogram
        t0 = time.perf_counter()
        get_variogram(space[:64], tt[:64], vv[:64], smax, tmax, ns, nt, None)      # JIT (same dtypes)
        jit_s = time.perf_counter() - t0
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            out = get_variogram(space, tt, vv, smax, tmax, ns, nt, None)
            times.append(time.perf_counter() - t0)
        return {"kind": "reference", "cores": 1, "pairs": pairs, "times": times, "jit_s": jit_s,
                "binned": float(np.nansum(out[1])),
                "what": "stif.utils.get_variogram (numba, unmodified; serial loop: 1 core)"}
    from . import variogram as ov
    ov.raw(space[:64], tt[:64], vv[:64], smax, tmax, ns, nt)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        h, c = ov.raw(space, tt, vv, smax, tmax, ns, nt)
        times.append(time.perf_counter() - t0)
    return {"kind": "port", "cores": 1, "pairs": pairs, "times": times, "jit_s": 0.0,
            "binned": float(c.sum()), "what": f"C restatement oracle/oracle_variogram.c ({why})"}


def kriging_sample(obs, targets, models, params, k, smax, tmax, batch_size=200):
    """get_kriging_prediction of the reference on a bounded target sample against ALL
    observations.  obs = (x, y, t, z), targets = (tx, ty, tt)."""
    stif, why = _load_reference()
    ox, oy, ot, oz = obs
    tx, ty, tt = targets
    n_t = len(tt)
    if stif is not None:
        import numba
        import pandas as pd
        df = pd.DataFrame({"x": ox, "y": oy, "time": ot, "z": oz})
        data = stif.Data(df, space_cols=["x", "y"], time_col="time", predictand_col="z", covariate_cols=[])
        p = stif.Predictor(data, None)
        p._prepare_geostatistics()
        # fixed parameter vector instead of the Nelder-Mead fit
        p._variogram_fit = types.SimpleNamespace(x=np.asarray(params, dtype=float))
        p._variogram_models = list(models)
        p._variogram_model_function = p._create_variogram_model_function()
        p._kriging_weights_function = p._create_kriging_weights_function()
        p._kriging_function = p._create_kriging_function()
        tsp = np.column_stack([tx, ty])
        t0 = time.perf_counter()
        p.get_kriging_prediction(tsp[:8], tt[:8], 1, k, smax, tmax, batch_size=8)      # JIT
        jit_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        mean, std = p.get_kriging_prediction(tsp, tt, 1, k, smax, tmax, batch_size=batch_size)
        dt = time.perf_counter() - t0
        return {"kind": "reference", "cores": int(numba.get_num_threads()), "n": n_t, "time": dt,
                "jit_s": jit_s, "mean": mean, "std": std,
                "what": "stif.Predictor.get_kriging_prediction (numba, unmodified; dense pre-filter on 1 "
                        f"core, nd_kriging on {numba.get_num_threads()} numba threads, batch_size={batch_size})"}
    from . import kriging as ok
    threads = os.cpu_count() or 1
    space, tsp = np.column_stack([ox, oy]), np.column_stack([tx, ty])
    ok.predict(space[:5000], ot[:5000], oz[:5000], tsp[:16], tt[:16], models, params, 1, k, smax, tmax)
    t0 = time.perf_counter()
    o = ok.predict(space, ot, oz, tsp, tt, models, params, 1, k, smax, tmax, n_threads=threads)
    dt = time.perf_counter() - t0
    return {"kind": "port", "cores": threads, "n": n_t, "time": dt, "jit_s": 0.0, "mean": o["mean"],
            "std": o["std"], "what": f"C restatement oracle/oracle_kriging.c, OpenMP over targets ({why})"}
