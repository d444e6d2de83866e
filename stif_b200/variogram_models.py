"""Parametric space-time variogram models -- HOST side (numpy), used only by the
scipy Nelder-Mead fit, which the north star keeps on the host.

Same formulas, names and parameter layouts as stif/variogram_models.py:7-218,
:311-398 of the reference; written vectorised over numpy arrays instead of numba
scalars.  The DEVICE evaluation of gamma(h, t) used by kriging lives in
csrc/kriging.cu (`gamma_st`); tests check the two agree.
"""
import numpy as np


def spherical(h, r, c0, b=0.0):
    """b + c0*(1.5 h/r - 0.5 (h/r)^3) for h <= r, else b + c0 (reference :40-44)."""
    h = np.asarray(h, dtype=np.float64)
    with np.errstate(all="ignore"):
        u = h / (r / 1.0)
        inside = b + c0 * ((1.5 * u) - (0.5 * (u * u * u)))
    return np.where(h <= r, inside, b + c0)


def gaussian(h, r, c0, b=0.0):
    """b + c0*(1 - exp(-h^2/(r/2)^2)) (reference :70-71)."""
    h = np.asarray(h, dtype=np.float64)
    a = r / 2.0
    with np.errstate(all="ignore"):
        return b + c0 * (1.0 - np.exp(-((h * h) / (a * a))))


def sum_model(h, t, x, model_space, model_time, model_metric):
    """x = [r_s, r_t, c0_s, c0_t, b_s, b_t] (reference :98-99)."""
    return model_space(h, x[0], x[2], x[4]) + model_time(t, x[1], x[3], x[5])


def product_model(h, t, x, model_space, model_time, model_metric):
    """x as sum_model (reference :126-127)."""
    return model_space(h, x[0], x[2], x[4]) * model_time(t, x[1], x[3], x[5])


def product_sum_model(h, t, x, model_space, model_time, model_metric):
    """x = [..6.., k]; k*prod + prod, exactly as the reference writes it (:155-157)."""
    p = product_model(h, t, x[:6], model_space, model_time, None)
    return x[6] * p + p


def metric_model(h, t, x, model_space, model_time, model_metric):
    """x = [r, c0, b, ani] (reference :184-186)."""
    h = np.asarray(h, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    d = np.sqrt(h * h + x[3] * x[3] * t * t)
    return model_metric(d, x[0], x[1], x[2])


def sum_metric_model(h, t, x, model_space, model_time, model_metric):
    """x = [r_s, r_t, r_m, c0_s, c0_t, c0_m, b_s, b_t, b_m, ani] (reference :213-218)."""
    h = np.asarray(h, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    d = np.sqrt(h * h + x[9] * x[9] * t * t)
    return (model_space(h, x[0], x[3], x[6]) + model_time(t, x[1], x[4], x[7])) \
        + model_metric(d, x[2], x[5], x[8])


variogram_model_dict = {
    "spherical": spherical,
    "gaussian": gaussian,
    "sum": sum_model,
    "product": product_model,
    "product_sum": product_sum_model,
    "metric": metric_model,
    "sum_metric": sum_metric_model,
}

N_PARAMS = {"sum": 6, "product": 6, "product_sum": 7, "metric": 4, "sum_metric": 10}


def prediction_grid(x, st_model, model_space, model_time, model_metric, bins_space, bins_time):
    """gamma on the (space-lag x time-lag) grid of bin centres (reference :222-261)."""
    h, t = np.meshgrid(np.asarray(bins_space, float), np.asarray(bins_time, float), indexing="ij")
    return np.asarray(st_model(h, t, x, model_space, model_time, model_metric), dtype=np.float64)


def weighted_mean_square_error(x, st_model, model_space, model_time, model_metric, bins_space,
                               bins_time, empirical_variogram, weights):
    """Objective of the Nelder-Mead fit (reference :264-308)."""
    err = empirical_variogram - prediction_grid(x, st_model, model_space, model_time, model_metric,
                                                bins_space, bins_time)
    return np.average(np.square(err), weights=weights)


def calc_weights(bins_space, bins_time, ani, samples_per_bin):
    """N(h,u) / (h^2 + (ani*u)^2), normalised to max 1 (reference :311-346)."""
    bs = np.asarray(bins_space, dtype=np.float64)[:, None]
    bt = np.asarray(bins_time, dtype=np.float64)[None, :]
    dist_sq = np.square(ani * bt) + np.square(bs)
    w = np.asarray(samples_per_bin).astype(float) / dist_sq
    return w / np.max(w)


def get_initial_parameters(model_str, empirical_variogram, space_dist_max, time_dist_max, ani):
    """Starting simplex centre per model (reference :349-398)."""
    ev = np.asarray(empirical_variogram)
    sill_s, sill_t, sill_m = ev[:, 0].max(), ev[0, :].max(), ev.max()
    half_nugget = ev.min() / 2
    r_s = r_m = space_dist_max
    r_t = time_dist_max
    if model_str == "sum":
        p = [r_s, r_t, sill_s, sill_t, half_nugget, half_nugget]
    elif model_str == "product":
        p = [r_s, r_t, np.sqrt(sill_s), np.sqrt(sill_t), np.sqrt(half_nugget), np.sqrt(half_nugget)]
    elif model_str == "product_sum":
        p = [r_s, r_t, np.sqrt(sill_s) / 2, np.sqrt(sill_t) / 2, np.sqrt(half_nugget / 2),
             np.sqrt(half_nugget / 2), 0.5]
    elif model_str == "metric":
        p = [r_m, sill_m, half_nugget * 2, ani]
    elif model_str == "sum_metric":
        p = [r_s, r_t, r_m, sill_s / 2, sill_t / 2, sill_m / 2, half_nugget / 2, half_nugget / 2,
             half_nugget, ani]
    else:
        raise KeyError(model_str)
    return np.array(p)
