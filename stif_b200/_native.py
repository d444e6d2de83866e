"""ctypes binding of libstif_b200.so (see include/stif_b200.h).

There is NO CPU fallback: if the shared library is missing or no B200 is present the
calls raise.  Nothing in this module imports `oracle/`.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstif_b200.so")

STIF_OK = 0
STIF_EINVAL = -1
STIF_ECUDA = -2
STIF_ENOMEM = -3
STIF_ESTATE = -4
STIF_ENONFINITE = -5

VARIO_PEEL = 1
VARIO_NO_CULL = 2
VARIO_SPACE_SORT = 4   # force the space-time sort mode (default: from 163 840 observations)
VARIO_TIME_SORT = 8    # forbid it
VARIO_DETERMINISTIC = 32  # bit-reproducible fixed-point sums (slower)
VARIO_COUNTS_F64 = 16  # variogram_dev: counts written as float64 (one fused all-reduce buffer)
KRIG_F32_STORAGE = 1
KRIG_FORCE_LU = 2
KRIG_TRY_CHOL = 4
KRIG_TRY_LDL = 8

ST_MODELS = {"sum": 0, "product": 1, "product_sum": 2, "metric": 3, "sum_metric": 4}
MARGINAL_MODELS = {"spherical": 0, "gaussian": 1}
ST_NPARAMS = {"sum": 6, "product": 6, "product_sum": 7, "metric": 4, "sum_metric": 10}

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_u64_p = ctypes.POINTER(ctypes.c_uint64)
_vp = ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/stif_b200.h declares
SIGNATURES = {
    "stif_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "stif_ctx_destroy": (ctypes.c_int, [_vp]),
    "stif_last_error": (ctypes.c_char_p, []),
    "stif_version": (ctypes.c_int, []),
    "stif_device_count": (ctypes.c_int, []),
    "stif_sm_count": (ctypes.c_int, [_vp]),
    "stif_variogram_dev": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_double,
                                          ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          _vp, _vp, _vp]),
    "stif_variogram_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_double,
                                           ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           _vp, _vp]),
    "stif_variogram_bound": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, _vp, _vp]),
    "stif_variogram_sampled_dev": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64,
                                                  ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                  ctypes.c_uint64, ctypes.c_uint64, _vp, _vp, _vp]),
    "stif_variogram_sampled_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64,
                                                   ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                   ctypes.c_uint64, ctypes.c_uint64, _vp, _vp]),
    "stif_variogram_stats": (ctypes.c_int, [_vp, _c_u64_p]),
    "stif_kriging_set_obs": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int]),
    "stif_kriging_pairs_count": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_double,
                                                ctypes.c_double, ctypes.POINTER(ctypes.c_int64)]),
    "stif_kriging_pairs_fetch": (ctypes.c_int, [_vp, _vp]),
    "stif_kriging_set_residuals": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int]),
    "stif_kriging_set_model": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, _vp, ctypes.c_int]),
    "stif_kriging_weights_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp]),
    "stif_variogram_model_eval": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, _vp]),
    "stif_fit_set_grid": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp]),
    "stif_fit_wmse": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp,
                                     ctypes.c_int, _c_double_p]),
    "stif_kriging_predict_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                 ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "stif_kriging_predict_dev": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "stif_debug_numpy_row_sum": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    "stif_kriging_stats": (ctypes.c_int, [_vp, _c_u64_p]),
    "stif_kriging_solver_stats": (ctypes.c_int, [_vp, _c_u64_p]),
    "stif_last_phase_ms": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "stif_bench_dfma": (ctypes.c_int, [_vp, ctypes.c_int, _c_double_p, ctypes.POINTER(ctypes.c_float)]),
}

_lib = None


class StifNativeError(RuntimeError):
    """CUDA/runtime failure inside libstif_b200."""


def load_library():
    """dlopen libstif_b200.so and declare every prototype. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StifNativeError(
            f"{LIB_PATH} is not built (run `python -m stif_b200.build`); "
            "stif_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def _raise(code, where):
    msg = load_library().stif_last_error().decode("utf-8", "replace")
    text = f"{where}: {msg} (code {code})"
    if code in (STIF_EINVAL, STIF_ESTATE):
        raise ValueError(text)
    if code == STIF_ENOMEM:
        raise MemoryError(text)
    if code == STIF_ENONFINITE:
        raise np.linalg.LinAlgError(text)
    raise StifNativeError(text)


def _check(code, where):
    if code != STIF_OK:
        _raise(code, where)


def _f64(a):
    """Contiguous float64 copy/view; int64 and bool convert exactly (|t| < 2^53)."""
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data


def split_space(space):
    """(n,2) C- or F-ordered array -> two contiguous float64 columns."""
    space = np.asarray(space)
    if space.ndim != 2 or space.shape[1] != 2:
        raise ValueError(f"space must have shape (n, 2), got {space.shape}")
    return _f64(space[:, 0]), _f64(space[:, 1])


class Context:
    """One libstif_b200 context = one CUDA device. Not thread-safe."""

    def __init__(self, device=-1):
        self._lib = load_library()
        h = _vp()
        _check(self._lib.stif_ctx_create(int(device), ctypes.byref(h)), "stif_ctx_create")
        self._h = h
        self._obs_id = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.stif_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self):
        return self._lib.stif_sm_count(self._h)

    # ---- variogram ---------------------------------------------------------------
    def variogram_host(self, x, y, t, v, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                       mode=0, mean=0.0, flags=VARIO_PEEL, part=0, nparts=1):
        x, y, t, v = _f64(x), _f64(y), _f64(t), _f64(v)
        n = len(x)
        if not (len(y) == n and len(t) == n and len(v) == n):
            raise ValueError("x, y, t, v must have the same length")
        nb = int(n_space_bins) * int(n_time_bins)
        counts = np.zeros(max(nb, 0), dtype=np.uint64)
        sums = np.zeros(max(nb, 0), dtype=np.float64)
        _check(self._lib.stif_variogram_host(
            self._h, _ptr(x), _ptr(y), _ptr(t), _ptr(v), n, float(space_dist_max),
            float(time_dist_max), int(n_space_bins), int(n_time_bins), int(mode), float(mean),
            int(flags), int(part), int(nparts), _ptr(counts), _ptr(sums)), "stif_variogram_host")
        shape = (int(n_space_bins), int(n_time_bins))
        return counts.reshape(shape), sums.reshape(shape)

    def variogram_bound(self, idx, space_dist_max, time_dist_max, n_space_bins, n_time_bins, mode=0, mean=0.0,
                        flags=VARIO_PEEL, part=0, nparts=1):
        """Variogram of the bound observation set (kriging_set_obs), restricted to `idx` (int64
        array) or all of it (idx=None): nothing but the index list travels to the device."""
        ii = None if idx is None else np.ascontiguousarray(idx, dtype=np.int64).reshape(-1)
        nb = int(n_space_bins) * int(n_time_bins)
        counts = np.zeros(max(nb, 0), dtype=np.uint64)
        sums = np.zeros(max(nb, 0), dtype=np.float64)
        _check(self._lib.stif_variogram_bound(
            self._h, _ptr(ii), 0 if ii is None else len(ii), float(space_dist_max), float(time_dist_max),
            int(n_space_bins), int(n_time_bins), int(mode), float(mean), int(flags), int(part), int(nparts),
            _ptr(counts), _ptr(sums)), "stif_variogram_bound")
        shape = (int(n_space_bins), int(n_time_bins))
        return counts.reshape(shape), sums.reshape(shape)

    def variogram_dev(self, x_ptr, y_ptr, t_ptr, v_ptr, n, space_dist_max, time_dist_max,
                      n_space_bins, n_time_bins, counts_ptr, sums_ptr, mode=0, mean=0.0,
                      flags=VARIO_PEEL, part=0, nparts=1, stream=0):
        """Device pointers (ints, e.g. torch tensor.data_ptr()); asynchronous on `stream`."""
        _check(self._lib.stif_variogram_dev(
            self._h, x_ptr, y_ptr, t_ptr, v_ptr, int(n), float(space_dist_max),
            float(time_dist_max), int(n_space_bins), int(n_time_bins), int(mode), float(mean),
            int(flags), int(part), int(nparts), counts_ptr, sums_ptr, stream or None),
            "stif_variogram_dev")

    def variogram_sampled_host(self, x, y, t, v, space_dist_max, time_dist_max, n_space_bins,
                               n_time_bins, n_samples, mode=0, mean=0.0, seed=None, seed_offset=0):
        """Sampled-pair branch (el_max).  seed=None draws a fresh seed (like the reference's
        OS-seeded RNG) unless STIF_B200_SEED is set in the environment; pass an int (or set the
        variable) for a reproducible call."""
        x, y, t, v = _f64(x), _f64(y), _f64(t), _f64(v)
        n = len(x)
        if seed is None:
            env = os.environ.get("STIF_B200_SEED")
            seed = int(env) if env else int.from_bytes(os.urandom(8), "little")
        seed = (int(seed) + 0x9E3779B97F4A7C15 * int(seed_offset)) & 0xFFFFFFFFFFFFFFFF
        nb = int(n_space_bins) * int(n_time_bins)
        counts = np.zeros(max(nb, 0), dtype=np.uint64)
        sums = np.zeros(max(nb, 0), dtype=np.float64)
        _check(self._lib.stif_variogram_sampled_host(
            self._h, _ptr(x), _ptr(y), _ptr(t), _ptr(v), n, float(space_dist_max),
            float(time_dist_max), int(n_space_bins), int(n_time_bins), int(mode), float(mean),
            int(n_samples), seed, _ptr(counts), _ptr(sums)), "stif_variogram_sampled_host")
        shape = (int(n_space_bins), int(n_time_bins))
        return counts.reshape(shape), sums.reshape(shape)

    def variogram_stats(self):
        out = (ctypes.c_uint64 * 4)()
        _check(self._lib.stif_variogram_stats(self._h, out), "stif_variogram_stats")
        return {"pairs": int(out[0]), "pairs_evaluated": int(out[1]), "pairs_in_window": int(out[2]),
                "items": int(out[3])}

    # ---- kriging -----------------------------------------------------------------
    def kriging_set_obs(self, x, y, t, resid):
        x, y, t, resid = _f64(x), _f64(y), _f64(t), _f64(resid)
        n = len(x)
        if not (len(y) == n and len(t) == n and len(resid) == n):
            raise ValueError("x, y, t, resid must have the same length")
        _check(self._lib.stif_kriging_set_obs(self._h, _ptr(x), _ptr(y), _ptr(t), _ptr(resid), n, 0),
               "stif_kriging_set_obs")

    def kriging_neighbour_pairs(self, tx, ty, tt, space_dist_max, time_dist_max, leave_out=None):
        """(P, 2) int64 [obs_idx, target_idx] of Data.get_kriging_idxs (stif/data.py:170-212)."""
        tx, ty, tt = _f64(tx), _f64(ty), _f64(tt)
        n = len(tx)
        lo = None if leave_out is None else np.ascontiguousarray(leave_out, dtype=np.int64)
        if lo is not None and len(lo) != n:
            raise ValueError("leave_out_idxs must have one entry per target")
        total = ctypes.c_int64()
        _check(self._lib.stif_kriging_pairs_count(self._h, _ptr(tx), _ptr(ty), _ptr(tt), _ptr(lo), n,
                                                  float(space_dist_max), float(time_dist_max),
                                                  ctypes.byref(total)), "stif_kriging_pairs_count")
        out = np.empty((total.value, 2), dtype=np.int64)
        if total.value:
            _check(self._lib.stif_kriging_pairs_fetch(self._h, _ptr(out)), "stif_kriging_pairs_fetch")
        return out

    def kriging_set_residuals(self, resid):
        r = _f64(resid)
        _check(self._lib.stif_kriging_set_residuals(self._h, _ptr(r), len(r), 0), "stif_kriging_set_residuals")

    def kriging_set_obs_dev(self, x_ptr, y_ptr, t_ptr, r_ptr, n):
        _check(self._lib.stif_kriging_set_obs(self._h, x_ptr, y_ptr, t_ptr, r_ptr, int(n), 1),
               "stif_kriging_set_obs")

    def kriging_set_model(self, st_model, space_model, time_model, metric_model, params):
        if st_model not in ST_MODELS:
            raise KeyError(st_model)
        for m in (space_model, time_model, metric_model):
            if m not in MARGINAL_MODELS:
                raise KeyError(m)
        p = _f64(params)
        _check(self._lib.stif_kriging_set_model(
            self._h, ST_MODELS[st_model], MARGINAL_MODELS[space_model], MARGINAL_MODELS[time_model],
            MARGINAL_MODELS[metric_model], _ptr(p), len(p)), "stif_kriging_set_model")

    def kriging_weights(self, xs, ys, ts, gamma_vec):
        """Ordinary-kriging weights (float64) of one neighbour set: calc_kriging_weights of the
        reference (stif/predictor.py:566-593), solved like np.linalg.lstsq on the device."""
        xs, ys, ts, g = _f64(xs).ravel(), _f64(ys).ravel(), _f64(ts).ravel(), _f64(gamma_vec).ravel()
        k = len(g)
        if not (len(xs) == k and len(ys) == k and len(ts) == k):
            raise ValueError("coordinates and kriging vector must have the same length")
        w = np.empty(k, dtype=np.float64)
        _check(self._lib.stif_kriging_weights_host(self._h, _ptr(xs), _ptr(ys), _ptr(ts), _ptr(g), k, _ptr(w)),
               "stif_kriging_weights")
        return w

    def variogram_model_eval(self, h, t):
        h_b, t_b = np.broadcast_arrays(np.asarray(h, dtype=np.float64), np.asarray(t, dtype=np.float64))
        shape = h_b.shape
        h_c, t_c = _f64(h_b).ravel(), _f64(t_b).ravel()
        out = np.empty(h_c.shape, dtype=np.float64)
        _check(self._lib.stif_variogram_model_eval(self._h, _ptr(h_c), _ptr(t_c), h_c.size, _ptr(out)),
               "stif_variogram_model_eval")
        return out.reshape(shape)

    # ---- variogram-model fit objective ---------------------------------------------
    def fit_set_grid(self, bins_space, bins_time, empirical, weights):
        bs, bt = _f64(bins_space).ravel(), _f64(bins_time).ravel()
        emp, w = _f64(empirical), _f64(weights)
        if emp.shape != (len(bs), len(bt)) or w.shape != emp.shape:
            raise ValueError("empirical / weights must have shape (len(bins_space), len(bins_time))")
        self._fit_keep = (bs, bt, emp, w)
        _check(self._lib.stif_fit_set_grid(self._h, _ptr(bs), len(bs), _ptr(bt), len(bt), _ptr(emp), _ptr(w)),
               "stif_fit_set_grid")

    def fit_wmse(self, st_model, space_model, time_model, metric_model, params):
        p = _f64(params).ravel()
        out = ctypes.c_double()
        _check(self._lib.stif_fit_wmse(self._h, ST_MODELS[st_model], MARGINAL_MODELS[space_model],
                                       MARGINAL_MODELS[time_model], MARGINAL_MODELS[metric_model],
                                       _ptr(p), len(p), ctypes.byref(out)), "stif_fit_wmse")
        return out.value

    def kriging_predict_host(self, tx, ty, tt, min_pts, max_pts, space_dist_max, time_dist_max,
                             leave_out=None, flags=KRIG_F32_STORAGE, want_weights=False,
                             out_mean=None, out_std=None):
        """out_mean / out_std: optional preallocated float64 arrays of n targets (e.g. pinned
        host memory, which makes the device-to-host copy of the result a plain DMA)."""
        tx, ty, tt = _f64(tx), _f64(ty), _f64(tt)
        n = len(tx)
        if not (len(ty) == n and len(tt) == n):
            raise ValueError("tx, ty, tt must have the same length")
        lo = None
        if leave_out is not None:
            lo = np.ascontiguousarray(leave_out, dtype=np.int64)
            if len(lo) != n:
                raise ValueError("leave_out_idxs must have one entry per target")
        def _out(a):
            if a is None:
                return np.empty(n, dtype=np.float64)
            if a.dtype != np.float64 or a.shape != (n,) or not a.flags.c_contiguous:
                raise ValueError("out_mean / out_std must be contiguous float64 arrays of n targets")
            return a
        mean, std = _out(out_mean), _out(out_std)
        w = g = idx = nused = None
        if want_weights:
            w = np.zeros((n, int(max_pts)), dtype=np.float32)
            g = np.zeros((n, int(max_pts)), dtype=np.float32)
            idx = np.zeros((n, int(max_pts)), dtype=np.uint32)
            nused = np.zeros(n, dtype=np.int32)
        _check(self._lib.stif_kriging_predict_host(
            self._h, _ptr(tx), _ptr(ty), _ptr(tt), _ptr(lo), n, int(min_pts), int(max_pts),
            float(space_dist_max), float(time_dist_max), int(flags), _ptr(mean), _ptr(std),
            _ptr(w), _ptr(g), _ptr(idx), _ptr(nused)), "stif_kriging_predict_host")
        if want_weights:
            return mean, std, w, g, idx, nused
        return mean, std

    def kriging_predict_dev(self, tx_ptr, ty_ptr, tt_ptr, n, min_pts, max_pts, space_dist_max,
                            time_dist_max, mean_ptr, std_ptr, leave_out_ptr=None,
                            flags=KRIG_F32_STORAGE, w_ptr=None, g_ptr=None, idx_ptr=None,
                            nused_ptr=None, stream=0):
        _check(self._lib.stif_kriging_predict_dev(
            self._h, tx_ptr, ty_ptr, tt_ptr, leave_out_ptr, int(n), int(min_pts), int(max_pts),
            float(space_dist_max), float(time_dist_max), int(flags), mean_ptr, std_ptr, w_ptr, g_ptr,
            idx_ptr, nused_ptr, stream or None), "stif_kriging_predict_dev")

    def kriging_stats(self):
        out = (ctypes.c_uint64 * 4)()
        _check(self._lib.stif_kriging_stats(self._h, out), "stif_kriging_stats")
        return {"candidates_examined": int(out[0]), "candidates_in_window": int(out[1]),
                "targets_below_min": int(out[2]), "targets_singular": int(out[3])}

    def kriging_solver_stats(self):
        out = (ctypes.c_uint64 * 2)()
        _check(self._lib.stif_kriging_solver_stats(self._h, out), "stif_kriging_solver_stats")
        return {"cholesky_to_lu": int(out[0]), "cholesky_enabled": bool(out[1] & 1),
                "ldl_enabled": bool(out[1] & 2), "first_pass_to_lu": int(out[0])}

    def debug_numpy_row_sum(self, rows):
        """np.sum(rows, axis=1) by the DEVICE summation code of the float32-storage epilogue."""
        rows = np.ascontiguousarray(rows)
        if rows.dtype not in (np.float32, np.float64) or rows.ndim != 2:
            raise ValueError("rows must be a 2-D float32 or float64 array")
        out = np.empty(rows.shape[0], dtype=rows.dtype)
        _check(self._lib.stif_debug_numpy_row_sum(self._h, _ptr(rows), rows.shape[0], rows.shape[1],
                                                  int(rows.dtype == np.float32), _ptr(out)),
               "stif_debug_numpy_row_sum")
        return out

    # ---- measurement -------------------------------------------------------------
    def last_phase_ms(self):
        out = (ctypes.c_float * 4)()
        _check(self._lib.stif_last_phase_ms(self._h, out), "stif_last_phase_ms")
        return [float(v) for v in out]

    def bench_dfma(self, iters=4096):
        rate = ctypes.c_double()
        ms = ctypes.c_float()
        _check(self._lib.stif_bench_dfma(self._h, int(iters), ctypes.byref(rate), ctypes.byref(ms)),
               "stif_bench_dfma")
        return rate.value, ms.value


def device_count():
    return int(load_library().stif_device_count())


_default_ctx = None


def default_context():
    """Process-wide context.

    * under torchrun / with STIF_B200_DEVICE set: ONE GPU (LOCAL_RANK / the given index) -- the
      ranks split the work between them (stif_b200/parallel.py);
    * otherwise: EVERY visible GPU of the box from this one process (MultiContext: one context and
      one host thread per device; CUDA_VISIBLE_DEVICES or STIF_B200_DEVICES="0,2,5" narrow the
      set).  With one visible GPU this is a plain Context."""
    global _default_ctx
    if _default_ctx is None:
        one = os.environ.get("STIF_B200_DEVICE", os.environ.get("LOCAL_RANK"))
        if one is not None or "WORLD_SIZE" in os.environ:
            _default_ctx = Context(int(one if one is not None else -1))
        else:
            env = os.environ.get("STIF_B200_DEVICES")
            devices = [int(d) for d in env.split(",")] if env else list(range(max(device_count(), 1)))
            if len(devices) > 1:
                from .multi import MultiContext
                _default_ctx = MultiContext(devices)
            else:
                _default_ctx = Context(devices[0])
    return _default_ctx
