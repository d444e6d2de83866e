"""Host helpers kept for API compatibility with stif/utils.py.

The hot functions of the reference's utils module -- get_variogram / get_covariogram
(stif/utils.py:104-279) -- are served by the CUDA library; these thin wrappers have the
reference's signatures and return tuples and call straight into it (no CPU fallback).
calc_distance_matrix_1d/2d (stif/utils.py:5-44) are fused into the kriging assembly
kernel on the device; the numpy versions here exist for callers that import them.
"""
import numpy as np

from . import _native, parallel


def calc_distance_matrix_1d(vec):
    """|vec_i - vec_j| (stif/utils.py:5-23); dtype follows the input."""
    vec = np.asarray(vec)
    return np.abs(vec[:, None] - vec[None, :])


def calc_distance_matrix_2d(vec):
    """Euclidean distance matrix of (n,2) points (stif/utils.py:26-44)."""
    vec = np.asarray(vec)
    d0 = vec[:, None, 0] - vec[None, :, 0]
    d1 = vec[:, None, 1] - vec[None, :, 1]
    return np.sqrt(d0 ** 2 + d1 ** 2)


def _finish(sums, counts, scale_fn):
    norm = counts.astype(np.float64)          # reference returns counts as float64
    with np.errstate(invalid="ignore", divide="ignore"):
        out = scale_fn(sums, norm)
    out[norm == 0] = np.nan                   # stif/utils.py:186-188
    return out, norm


def _run(space, time, val, space_dist_max, time_dist_max, n_space_bins, n_time_bins, n_samples,
         mode, ctx=None):
    ctx = ctx or _native.default_context()
    x, y = _native.split_space(space)
    t = _native._f64(time)
    v = _native._f64(val)
    mean = float(np.mean(v)) if (mode == 1 and len(v)) else 0.0
    if n_samples is None:
        counts, sums = parallel.variogram_all_ranks(
            ctx, x, y, t, v, space_dist_max, time_dist_max, n_space_bins, n_time_bins, mode, mean)
    else:
        counts, sums = parallel.variogram_sampled_all_ranks(
            ctx, x, y, t, v, space_dist_max, time_dist_max, n_space_bins, n_time_bins, mode, mean,
            int(n_samples))
    ws = space_dist_max / n_space_bins        # stif/utils.py:146-147
    wt = time_dist_max / n_time_bins
    return counts, sums, ws, wt, v


def get_variogram(space, time, val, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                  n_samples=None, ctx=None):
    """Drop-in for stif.utils.get_variogram (stif/utils.py:104-190)."""
    counts, sums, ws, wt, _ = _run(space, time, val, space_dist_max, time_dist_max, n_space_bins,
                                   n_time_bins, n_samples, 0, ctx)
    vg, norm = _finish(sums, counts, lambda s, c: (s * 0.5) / c)
    return vg, norm, ws, wt


def get_covariogram(features, time, val, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                    n_samples=None, ctx=None):
    """Drop-in for stif.utils.get_covariogram (stif/utils.py:193-279)."""
    counts, sums, ws, wt, v = _run(features, time, val, space_dist_max, time_dist_max, n_space_bins,
                                   n_time_bins, n_samples, 1, ctx)
    var = float(np.var(v)) if len(v) else np.nan
    cg, norm = _finish(sums, counts, lambda s, c: (s / c) / var)
    return cg, norm, ws, wt
