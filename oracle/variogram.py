"""Python face of the C variogram oracle (oracle_variogram.c) -- TEST INFRASTRUCTURE.

`get_variogram` / `get_covariogram` take the same arguments and return the same tuple
as stif.utils.get_variogram (stif/utils.py:104-190) / get_covariogram (:193-279) with
n_samples=None (all pairs).
"""
import ctypes

import numpy as np

from . import build as _build

FLAG_PEEL = 1
FLAG_ACC_FMA = 2
FLAG_D2_FMA_X = 4
FLAG_D2_UNFUSED = 8
FLAG_TRUE_DIV = 16
REFERENCE_FLAGS = FLAG_PEEL  # the executed numba codegen

_dp = ctypes.POINTER(ctypes.c_double)
_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_build.build())
        lib.oracle_variogram_raw.restype = ctypes.c_int64
        lib.oracle_variogram_raw.argtypes = [
            _dp, _dp, _dp, _dp, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int64,
            ctypes.c_int64, _dp, _dp]
        lib.oracle_variogram.restype = None
        lib.oracle_variogram.argtypes = [
            _dp, _dp, _dp, _dp, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, _dp, _dp,
            _dp]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _soa(space, time, val):
    space = np.asarray(space)
    x = np.ascontiguousarray(space[:, 0], dtype=np.float64)
    y = np.ascontiguousarray(space[:, 1], dtype=np.float64)
    t = np.ascontiguousarray(time, dtype=np.float64)
    v = np.ascontiguousarray(val, dtype=np.float64)
    return x, y, t, v


def raw(space, time, val, space_dist_max, time_dist_max, n_space_bins, n_time_bins, mode=0,
        flags=REFERENCE_FLAGS, i_begin=0, i_end=None):
    """Raw accumulators (sums, counts) of the pair loop; rows i in [i_begin, i_end)."""
    lib = _load()
    x, y, t, v = _soa(space, time, val)
    n = len(x)
    hist = np.zeros((n_space_bins, n_time_bins))
    norm = np.zeros((n_space_bins, n_time_bins))
    mean = float(np.mean(v)) if (mode == 1 and n) else 0.0
    lib.oracle_variogram_raw(_p(x), _p(y), _p(t), _p(v), n, float(space_dist_max),
                             float(time_dist_max), int(n_space_bins), int(n_time_bins), int(mode),
                             mean, int(flags), int(i_begin), int(n if i_end is None else i_end),
                             _p(hist), _p(norm))
    return hist, norm


def _full(space, time, val, space_dist_max, time_dist_max, n_space_bins, n_time_bins, mode, flags):
    lib = _load()
    x, y, t, v = _soa(space, time, val)
    n = len(x)
    vg = np.zeros((n_space_bins, n_time_bins))
    norm = np.zeros((n_space_bins, n_time_bins))
    widths = np.zeros(2)
    mean = float(np.mean(v)) if mode == 1 else 0.0
    var = float(np.var(v)) if mode == 1 else 1.0
    lib.oracle_variogram(_p(x), _p(y), _p(t), _p(v), n, float(space_dist_max), float(time_dist_max),
                         int(n_space_bins), int(n_time_bins), int(mode), mean, var, int(flags),
                         _p(vg), _p(norm), _p(widths))
    return vg, norm, float(widths[0]), float(widths[1])


def get_variogram(space, time, val, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                  n_samples=None, flags=REFERENCE_FLAGS):
    if n_samples is not None:
        raise NotImplementedError("oracle covers the all-pairs branch (utils.py:90-92)")
    return _full(space, time, val, space_dist_max, time_dist_max, n_space_bins, n_time_bins, 0, flags)


def get_covariogram(space, time, val, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                    n_samples=None, flags=REFERENCE_FLAGS):
    if n_samples is not None:
        raise NotImplementedError("oracle covers the all-pairs branch (utils.py:90-92)")
    return _full(space, time, val, space_dist_max, time_dist_max, n_space_bins, n_time_bins, 1, flags)
