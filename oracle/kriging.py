"""Python face of the C kriging oracle (oracle_kriging.c) -- TEST INFRASTRUCTURE.

`predict(...)` restates Predictor.get_kriging_prediction / _get_kriging_weights
(stif/predictor.py:777-906) for a given observation set and fitted-model parameters.
The least-squares solve is LAPACK dgelsd from scipy's bundled OpenBLAS -- the routine the
reference reaches through numba's np.linalg.lstsq -- located at run time and handed to the
C code as a function pointer.
"""
import ctypes
import glob
import os

import numpy as np

from . import build as _build

ST_MODELS = {"sum": 0, "product": 1, "product_sum": 2, "metric": 3, "sum_metric": 4}
MARGINALS = {"spherical": 0, "gaussian": 1}
F32_STORAGE = 1

_dp = ctypes.POINTER(ctypes.c_double)


class _Model(ctypes.Structure):
    _fields_ = [("st", ctypes.c_int), ("ks", ctypes.c_int), ("kt", ctypes.c_int),
                ("km", ctypes.c_int), ("x", ctypes.c_double * 10)]


_lib = None
_dgelsd = None


def _find_dgelsd():
    """Address of dgelsd in scipy's OpenBLAS (the library numba's lstsq ends up in)."""
    global _dgelsd
    if _dgelsd is not None:
        return _dgelsd
    import scipy
    libs = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    for path in sorted(glob.glob(os.path.join(libs, "libscipy_openblas*.so*"))):
        lib = ctypes.CDLL(path)
        for name in ("scipy_dgelsd_", "dgelsd_"):
            try:
                fn = getattr(lib, name)
            except AttributeError:
                continue
            _dgelsd = (lib, ctypes.cast(fn, ctypes.c_void_p))
            return _dgelsd
    raise RuntimeError("LAPACK dgelsd not found in scipy.libs (needed by the kriging oracle)")


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_build.build())
        vp = ctypes.c_void_p
        lib.oracle_kriging.restype = ctypes.c_int
        lib.oracle_kriging.argtypes = [
            vp, vp, vp, vp, ctypes.c_int64, vp, vp, vp, vp, ctypes.c_int64, ctypes.c_int,
            ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.POINTER(_Model), vp, ctypes.c_int,
            ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.oracle_gamma_vec.restype = None
        lib.oracle_gamma_vec.argtypes = [ctypes.POINTER(_Model), vp, vp, ctypes.c_int64, vp]
        lib.oracle_kriging_weights.restype = ctypes.c_int
        lib.oracle_kriging_weights.argtypes = [vp, vp, vp, vp, ctypes.c_int, ctypes.POINTER(_Model),
                                               vp, vp, ctypes.POINTER(ctypes.c_int)]
        _lib = lib
    return _lib


def make_model(st_model, space_model, time_model, metric_model, params):
    m = _Model()
    m.st = ST_MODELS[st_model]
    m.ks = MARGINALS[space_model]
    m.kt = MARGINALS[time_model]
    m.km = MARGINALS[metric_model]
    p = np.asarray(params, dtype=np.float64)
    for i in range(10):
        m.x[i] = float(p[i]) if i < len(p) else 0.0
    return m


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def gamma(models, params, h, t):
    """gamma(h, t) of the fitted model, elementwise (stif/predictor.py:547-556)."""
    lib = _load()
    m = make_model(*models, params)
    hb, tb = np.broadcast_arrays(np.asarray(h, dtype=np.float64), np.asarray(t, dtype=np.float64))
    hc, tc = _f64(hb).ravel(), _f64(tb).ravel()
    out = np.empty_like(hc)
    lib.oracle_gamma_vec(ctypes.byref(m), hc.ctypes.data, tc.ctypes.data, hc.size, out.ctypes.data)
    return out.reshape(hb.shape)


def predict(obs_space, obs_time, residuals, space, time, models, params, min_kriging_points=1,
            max_kriging_points=10, space_dist_max=1.0, time_dist_max=1.0, leave_out_idxs=None,
            f32_storage=True, n_threads=0):
    """Returns a dict: mean, std (float64), w, gamma (float32 [n,k]), idx (uint32 [n,k]),
    n_used, n_cand (int32), tie (uint8), rank (int32)."""
    lib = _load()
    _, fn = _find_dgelsd()
    obs_space = np.asarray(obs_space)
    ox, oy = _f64(obs_space[:, 0]), _f64(obs_space[:, 1])
    ot, res = _f64(obs_time), _f64(residuals)
    space = np.asarray(space)
    tx, ty, tt = _f64(space[:, 0]), _f64(space[:, 1]), _f64(time)
    n, k = len(tt), int(max_kriging_points)
    lo = None if leave_out_idxs is None else np.ascontiguousarray(leave_out_idxs, dtype=np.int64)
    m = make_model(*models, params)
    out = {
        "mean": np.zeros(n), "std": np.zeros(n),
        "w": np.zeros((n, k), np.float32), "gamma": np.zeros((n, k), np.float32),
        "idx": np.zeros((n, k), np.uint32), "n_used": np.zeros(n, np.int32),
        "n_cand": np.zeros(n, np.int32), "tie": np.zeros(n, np.uint8), "rank": np.zeros(n, np.int32),
    }
    rc = lib.oracle_kriging(
        ox.ctypes.data, oy.ctypes.data, ot.ctypes.data, res.ctypes.data, len(ox), tx.ctypes.data,
        ty.ctypes.data, tt.ctypes.data, None if lo is None else lo.ctypes.data, n,
        int(min_kriging_points), k, float(space_dist_max), float(time_dist_max), ctypes.byref(m), fn,
        F32_STORAGE if f32_storage else 0, int(n_threads), out["mean"].ctypes.data,
        out["std"].ctypes.data, out["w"].ctypes.data, out["gamma"].ctypes.data,
        out["idx"].ctypes.data, out["n_used"].ctypes.data, out["n_cand"].ctypes.data,
        out["tie"].ctypes.data, out["rank"].ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_kriging failed (code {rc})")
    return out


def weights(xs, ys, ts, gvec, models, params):
    """Float64 ordinary-kriging weights of one neighbour set (predictor.py:566-593)."""
    lib = _load()
    _, fn = _find_dgelsd()
    xs, ys, ts, gvec = _f64(xs), _f64(ys), _f64(ts), _f64(gvec)
    k = len(xs)
    w = np.zeros(k)
    rank = ctypes.c_int()
    m = make_model(*models, params)
    rc = lib.oracle_kriging_weights(xs.ctypes.data, ys.ctypes.data, ts.ctypes.data, gvec.ctypes.data,
                                    k, ctypes.byref(m), fn, w.ctypes.data, ctypes.byref(rank))
    if rc != 0:
        raise RuntimeError(f"dgelsd failed (info {rc})")
    return w, rank.value


def kriging_idxs(obs_space, obs_time, space, time, space_dist_max, time_dist_max, leave_out_idxs=None):
    """Restatement of Data.get_kriging_idxs (stif/data.py:170-212) -- TEST INFRASTRUCTURE: the
    dense n_obs x n_targets filter with scipy's cdist, exactly as the reference writes it."""
    import scipy.spatial.distance
    obs_time = np.asarray(obs_time)
    time = np.asarray(time)
    space_dist_sq = scipy.spatial.distance.cdist(obs_space, space, metric="sqeuclidean")
    time_dist = scipy.spatial.distance.cdist(obs_time.reshape(-1, 1), time.reshape(-1, 1))
    index_mask = np.ones([len(obs_time), len(time)], dtype=bool)
    if leave_out_idxs is not None:
        index_mask[leave_out_idxs, np.arange(len(time))] = False
    is_close_enough = ((space_dist_sq < space_dist_max ** 2) & (time_dist < time_dist_max) &
                       (obs_time.reshape(-1, 1) <= time.reshape(1, -1)) & index_mask)
    return np.column_stack(np.nonzero(is_close_enough))
