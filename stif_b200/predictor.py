"""`Predictor`: space-time regression kriging with the reference's public API
(stif/predictor.py:42-917) and its two data-parallel kernels served by libstif_b200.so.

What runs where
  host (unchanged responsibilities, by the north star): covariate regression, the scipy
      Nelder-Mead variogram-model fit, cross-validation orchestration, .npz save/load.
  B200 (this package's C ABI, no CPU fallback): calc_empirical_variogram /
      calc_empirical_covariogram (stif/predictor.py:398-408, :471-481) and everything
      under get_kriging_prediction / _get_kriging_weights (:777-906): neighbour search,
      gamma-vector, top-k, system assembly + solve, mean/std epilogue.

Differences a user can observe (all opt-in or documented in DESIGN.md):
  * no numba JIT pause after fit_variogram_model (the device gamma(h,t) is parameterised);
  * `batch_size` is accepted and ignored (the device path never materialises n_obs x batch);
  * max_kriging_points <= 128;
  * ties at the top-k boundary are broken by lowest observation index (the reference's
    unstable argsort is not reproducible);
  * plotting helpers are not provided (matplotlib is out of scope for the hot path).
"""
import pickle
import types
import warnings

import numpy as np
import sklearn.metrics  # noqa: F401  (parity with the reference's import surface)
import sklearn.model_selection
from scipy.optimize import minimize

from . import _native, parallel
from .data import Data
from .utils import get_covariogram, get_variogram
from .variogram_models import (calc_weights, get_initial_parameters, variogram_model_dict,
                               weighted_mean_square_error)


def _bin_centres(n_bins, width):
    edges = np.arange(n_bins + 1) * width       # stif/predictor.py:410-414
    return (edges[:-1] + edges[1:]) / 2


class Predictor:
    """Predictor(data, covariate_model=None, cv_splits=5, resampling=None)."""

    def __init__(self, data: Data, covariate_model=None, cv_splits: int = 5, resampling=None,
                 emulate_f32_storage: bool = True, context=None, device_fit_objective: bool = True):
        self._data = data
        self._cv_splits = cv_splits
        self._resampling = resampling
        self._cov_model = covariate_model
        self._X = self._data.get_training_covariates()
        self._y = self._data.predictand

        self._cross_val_res = None
        self._variogram = None
        self._variogram_bins_space = None
        self._variogram_bins_time = None
        self._variogram_samples_per_bin = None
        self._variogram_fit = None
        self._variogram_models = None
        # the three callables fit_variogram_model installs in the reference
        # (stif/predictor.py:759-764); here they are bound methods over the device model
        self._variogram_model_function = None
        self._kriging_weights_function = None
        self._kriging_function = None

        self._is_binary = self._data.predictand.dtype == bool
        self._is_keras_model = self._cov_model.__class__.__module__ == "keras.src.engine.sequential"
        self._residuals = None

        # stock reference rounds weights / gamma-vector to float32 before the final sums
        # (stif/predictor.py:610-618); False keeps everything float64
        self._emulate_f32_storage = emulate_f32_storage
        # objective of the Nelder-Mead fit (prediction grid + weighted MSE) on the device
        # (stif/variogram_models.py:221-308); the optimiser itself stays scipy on the host
        self._device_fit_objective = device_fit_objective
        self._ctx = context
        self._obs_token = None
        # self._residuals_version (see __setattr__) is bumped whenever self._residuals is assigned:
        # the device copy is stale iff the counter moved (object identity is no key: CPython
        # reuses the address of a freed array)

    # ---- device plumbing ---------------------------------------------------------
    @property
    def context(self):
        if self._ctx is None:
            self._ctx = _native.default_context()
        return self._ctx

    def _upload_observations(self):
        """Replicate (x, y, t, residual) of the training data on this rank's GPU."""
        time = self._data.time_coords
        token = (self._residuals_version, len(time))
        mine = getattr(self.context, "_obs_owner", None) is self
        if self._obs_token == token and mine:
            return
        resid = np.asarray(self._residuals, dtype=np.float64)
        if mine and self._obs_token is not None and self._obs_token[1] == len(time):
            # same observation set, new residuals (cross-validation refits the covariate model
            # per fold): keep the coordinates and the space-time grid on the device
            self.context.kriging_set_residuals(resid)
        else:
            x, y = _native.split_space(self._data.space_coords)
            self.context.kriging_set_obs(x, y, time, resid)
        self.context._obs_owner = self
        self._obs_token = token

    def _push_model(self):
        st, sp, tm, mm = self._variogram_models
        self.context.kriging_set_model(st, sp, tm, mm, np.asarray(self._variogram_fit.x, dtype=np.float64))
        self.context._model_owner = self

    # ---- covariate regression (host) -----------------------------------------------
    def _prepare_geostatistics(self):
        self._residuals = self.get_residuals()

    def __setattr__(self, name, value):
        # every assignment of the residuals (here, by a subclass or by a test poking the
        # attribute like the reference's own tests do) invalidates the device copy
        if name == "_residuals":
            object.__setattr__(self, "_residuals_version", getattr(self, "_residuals_version", 0) + 1)
        object.__setattr__(self, name, value)

    def fit_covariate_model(self, train_idxs=slice(None)):
        X, y = self._X[train_idxs, :], self._y[train_idxs]
        if self._resampling is not None:
            X, y = self._resampling.fit_resample(X, y)
        if self._cov_model is not None:
            self._cov_model.fit(X, y)

    def save_covariate_model(self, filename):
        if self._is_keras_model:
            self._cov_model.save(filename)
        else:
            with open(filename, "wb") as fh:
                pickle.dump(self._cov_model, fh)

    def load_covariate_model(self, filename):
        if self._is_keras_model:
            try:
                from tensorflow import keras
            except ImportError as exc:
                raise ImportError("Keras is not installed. Please install Keras to load a Keras "
                                  "model.") from exc
            self._cov_model = keras.models.load_model(filename)
        else:
            with open(filename, "rb") as fh:
                self._cov_model = pickle.load(fh)
        self._prepare_geostatistics()

    @property
    def _covariate_prediction_function(self):
        if self._is_binary and not self._is_keras_model:
            return lambda X: self._cov_model.predict_proba(X)[:, 1]
        return self._cov_model.predict

    def get_covariate_prediction(self, idxs=slice(None)):
        if self._cov_model is None:
            return np.zeros(len(idxs))
        return self._covariate_prediction_function(self._X[idxs]).flatten()

    def calc_covariate_prediction(self, df):
        if self._cov_model is None:
            return np.zeros(len(df))
        return self._covariate_prediction_function(self._data.prepare_covariates(df))

    def get_residuals(self, idxs=slice(None)):
        if self._cov_model is not None:
            return self._y[idxs] - self.get_covariate_prediction(idxs)
        return self._y[idxs]

    # ---- cross validation (host orchestration; hot path on the GPU) ----------------
    def calc_cross_validation(self, kriging: bool = False, geostat_params: dict = dict(),
                              max_test_samples: int = -1, verbose: bool = False,
                              empirical_variogram_path=None):
        splitter = sklearn.model_selection.TimeSeriesSplit(n_splits=self._cv_splits)
        truths, predictions = [], []
        for fold, (train, test) in enumerate(splitter.split(self._X, self._y)):
            if verbose:
                print("Fold {}".format(fold))
                print("\t train: {} samples".format(len(train)))
                print("\t test: {} samples".format(len(test)))
            if max_test_samples > 0:
                test = np.random.choice(test, size=min(max_test_samples, len(test)), replace=False)
            self.fit_covariate_model(train)
            truths.append(self._y[test])
            prediction = self.get_covariate_prediction(test)
            if kriging:
                if empirical_variogram_path is None:
                    self.calc_empirical_variogram(train, **geostat_params.get("variogram_params", {}))
                else:
                    self.load_empirical_variogram(empirical_variogram_path)
                self.fit_variogram_model(**geostat_params.get("variogram_model_params", {}))
                kriging_mean, _ = self.get_kriging_prediction(
                    self._data.space_coords[test, :], self._data.time_coords[test],
                    leave_out_idxs=test, **geostat_params.get("kriging_params", {}))
                prediction = prediction + kriging_mean
            predictions.append(prediction)
        self._cross_val_res = truths, predictions

    def get_cross_val_metric(self, metric):
        if self._cross_val_res is None:
            raise ValueError("Calc cross validation first.")
        truths, predictions = self._cross_val_res
        return [metric(truths[i], predictions[i]) for i in range(self._cv_splits)]

    # ---- empirical variogram / covariogram (GPU) ------------------------------------
    def _variogram_inputs(self, idxs):
        space = self._data.space_coords[idxs, :]
        time = self._data.time_coords[idxs]
        self._prepare_geostatistics()
        if self._residuals is None and self._cov_model is not None:
            raise ValueError("Covariate model is set, but was not fitted yet. Please fit the "
                             "covariate model first, so the variogram can be calculated on the "
                             "residuals")
        return space, time, self._residuals[idxs]

    def _device_variogram(self, idxs, space_dist_max, time_dist_max, n_space_bins, n_time_bins, mode):
        """All-pairs variogram / covariogram of the training observations `idxs` from the
        DEVICE-RESIDENT observation set: coordinates are uploaded once per Predictor, residuals
        when they change (cross-validation refits the covariate model per fold), and a subset
        travels as its index list only (stif_variogram_bound).  Same arithmetic and result as
        utils.get_variogram / get_covariogram on the sliced arrays."""
        self._prepare_geostatistics()
        if self._residuals is None and self._cov_model is not None:
            raise ValueError("Covariate model is set, but was not fitted yet. Please fit the "
                             "covariate model first, so the variogram can be calculated on the "
                             "residuals")
        self._upload_observations()
        n = len(self._data.time_coords)
        whole = isinstance(idxs, slice) and idxs == slice(None)
        idx = None if whole else np.arange(n, dtype=np.int64)[idxs]
        resid = np.asarray(self._residuals, dtype=np.float64)
        vals = resid if whole else resid[idx]
        mean = float(np.mean(vals)) if (mode == 1 and len(vals)) else 0.0
        ctx = self.context
        rank, ws = parallel.world()
        kw = dict(mode=mode, mean=mean)
        if ws > 1:
            counts, sums = ctx.variogram_bound(idx, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                                               part=rank, nparts=ws, **kw)
            counts, sums = parallel.allreduce_histograms(counts, sums)
        elif hasattr(ctx, "contexts"):          # MultiContext: every visible GPU takes a partition
            counts, sums = ctx.variogram_bound(idx, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                                               n_obs=n, **kw)
        else:
            counts, sums = ctx.variogram_bound(idx, space_dist_max, time_dist_max, n_space_bins, n_time_bins, **kw)
        norm = counts.astype(np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            if mode == 0:
                out = (sums * 0.5) / norm                       # stif/utils.py:178-188
            else:
                out = (sums / norm) / (float(np.var(vals)) if len(vals) else np.nan)   # :267-272
        out[norm == 0] = np.nan
        return out, norm, space_dist_max / n_space_bins, time_dist_max / n_time_bins

    def calc_empirical_variogram(self, idxs=slice(None), space_dist_max=3, time_dist_max=10,
                                 n_space_bins=10, n_time_bins=10, el_max=None):
        if el_max is None:
            variogram, samples_per_bin, width_space, width_time = self._device_variogram(
                idxs, space_dist_max, time_dist_max, n_space_bins, n_time_bins, 0)
        else:
            space, time, residuals = self._variogram_inputs(idxs)
            variogram, samples_per_bin, width_space, width_time = get_variogram(
                space, time, residuals, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                el_max, ctx=self.context)
        self._variogram = variogram
        self._variogram_bins_space = _bin_centres(n_space_bins, width_space)
        self._variogram_bins_time = _bin_centres(n_time_bins, width_time)
        self._variogram_samples_per_bin = samples_per_bin

    def calc_empirical_covariogram(self, idxs=slice(None), space_dist_max=3, time_dist_max=10,
                                   n_space_bins=10, n_time_bins=10, el_max=None):
        if el_max is None:
            covariogram, samples_per_bin, width_space, width_time = self._device_variogram(
                idxs, space_dist_max, time_dist_max, n_space_bins, n_time_bins, 1)
        else:
            space, time, residuals = self._variogram_inputs(idxs)
            covariogram, samples_per_bin, width_space, width_time = get_covariogram(
                space, time, residuals, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                el_max, ctx=self.context)
        self._covariogram = covariogram
        self._covariogram_bins_space = _bin_centres(n_space_bins, width_space)
        self._covariogram_bins_time = _bin_centres(n_time_bins, width_time)
        self._covariogram_samples_per_bin = samples_per_bin

    def save_empirical_variogram(self, filename):
        if self._variogram is None:
            raise ValueError("Calc empirical variogoram first.")
        np.savez(filename, variogram=self._variogram, bins_space=self._variogram_bins_space,
                 bins_time=self._variogram_bins_time, samples_per_bin=self._variogram_samples_per_bin)

    def load_empirical_variogram(self, filename):
        with np.load(filename) as stored:
            self._variogram = stored["variogram"]
            self._variogram_bins_space = stored["bins_space"]
            self._variogram_bins_time = stored["bins_time"]
            self._variogram_samples_per_bin = stored["samples_per_bin"]

    # ---- variogram model (host fit, device evaluation) -------------------------------
    def _create_variogram_model_function(self):
        if self._variogram_fit is None:
            raise ValueError("Fit variogram model first.")
        self._push_model()

        def variogram_model(h, t):
            if getattr(self.context, "_model_owner", None) is not self:
                self._push_model()
            return self.context.variogram_model_eval(h, t)
        return variogram_model

    def _create_kriging_weights_function(self):
        if self._variogram_model_function is None:
            raise ValueError("Create variogram model function first.")
        return types.MethodType(Predictor._device_kriging_weights, self)

    def _create_kriging_function(self):
        if self._kriging_weights_function is None:
            raise ValueError("Create Kriging weights function first.")
        return types.MethodType(Predictor._device_kriging, self)

    def set_variogram_model(self, params, st_model="sum_metric", space_model="spherical",
                            time_model="spherical", metric_model="spherical"):
        """Install a variogram model with GIVEN parameters (no Nelder-Mead); what benchmarks
        and cross-validation with a precomputed fit use."""
        for name in (st_model, space_model, time_model, metric_model):
            if name not in variogram_model_dict:
                raise KeyError(name)
        self._variogram_fit = types.SimpleNamespace(x=np.asarray(params, dtype=np.float64), fun=None)
        self._variogram_models = [st_model, space_model, time_model, metric_model]
        self._variogram_model_function = self._create_variogram_model_function()
        self._kriging_weights_function = self._create_kriging_weights_function()
        self._kriging_function = self._create_kriging_function()

    def fit_variogram_model(self, st_model="sum_metric", space_model="spherical",
                            time_model="spherical", metric_model="spherical",
                            plot_anisotropy=False):
        if self._variogram is None:
            raise ValueError("Calc empirical variogoram first.")
        poly = np.polynomial.polynomial.polyfit
        slope_space = poly(self._variogram_bins_space, self._variogram[:, 0], deg=1)[1]
        slope_time = poly(self._variogram_bins_time, self._variogram[0, :], deg=1)[1]
        ani = slope_time / slope_space
        if plot_anisotropy:
            raise NotImplementedError("plotting is outside the scope of stif_b200")
        weights = calc_weights(self._variogram_bins_space, self._variogram_bins_time, ani,
                               self._variogram_samples_per_bin)
        x0 = get_initial_parameters(st_model, self._variogram, self._variogram_bins_space[-1],
                                    self._variogram_bins_time[-1], ani)
        if self._device_fit_objective:
            ctx = self.context
            ctx.fit_set_grid(self._variogram_bins_space, self._variogram_bins_time, self._variogram, weights)
            fit = minimize(lambda x: ctx.fit_wmse(st_model, space_model, time_model, metric_model, x), x0,
                           method="Nelder-Mead", options={"maxiter": 10000})
        else:
            fit = minimize(
                weighted_mean_square_error, x0,
                args=(variogram_model_dict[st_model], variogram_model_dict[space_model],
                      variogram_model_dict[time_model], variogram_model_dict[metric_model],
                      self._variogram_bins_space, self._variogram_bins_time, self._variogram, weights),
                method="Nelder-Mead", options={"maxiter": 10000})
        self._variogram_fit = fit
        self._variogram_models = [st_model, space_model, time_model, metric_model]
        self._variogram_model_function = self._create_variogram_model_function()
        self._kriging_weights_function = self._create_kriging_weights_function()
        self._kriging_function = self._create_kriging_function()

    def _get_variogram_model_grid(self):
        if self._variogram_fit is None:
            raise ValueError("Fit variogram model first.")
        h, t = np.meshgrid(self._variogram_bins_space, self._variogram_bins_time, indexing="ij")
        return self._variogram_model_function(h, t)

    # ---- kriging (GPU) ------------------------------------------------------------------
    def _device_kriging(self, space, time, min_kriging_points, max_kriging_points, space_dist_max,
                        time_dist_max, leave_out_idxs=None, want_weights=True):
        """Whole prediction pipeline of the reference's pre-filter + nd_kriging + epilogue for
        this rank's shard of the targets; gathers over ranks."""
        if self._residuals is None:
            self._prepare_geostatistics()
        self._upload_observations()
        if getattr(self.context, "_model_owner", None) is not self:
            self._push_model()
        space = np.asarray(space)
        time = np.asarray(time)
        tx, ty = _native.split_space(space.reshape(-1, 2))
        tt = _native._f64(time).reshape(-1)
        lo = None if leave_out_idxs is None else np.asarray(leave_out_idxs, dtype=np.int64).reshape(-1)
        flags = _native.KRIG_F32_STORAGE if self._emulate_f32_storage else 0

        def run(b, e):
            out = self.context.kriging_predict_host(
                tx[b:e], ty[b:e], tt[b:e], min_kriging_points, max_kriging_points, space_dist_max,
                time_dist_max, leave_out=None if lo is None else lo[b:e], flags=flags,
                want_weights=want_weights)
            n_def = self.context.kriging_stats()["targets_singular"]
            if n_def:
                warnings.warn(f"{n_def} kriging system(s) were rank-deficient (duplicate or co-located "
                              "observations); they received the minimum-norm solution, like the "
                              "reference's np.linalg.lstsq", RuntimeWarning, stacklevel=3)
            return out
        return parallel.kriging_all_ranks(run, len(tt))

    def _device_kriging_weights(self, kriging_vector, coords_spatial, coords_temporal):
        """calc_kriging_weights (stif/predictor.py:566-593): ordinary-kriging weights of ONE
        neighbour set from its gamma-vector and coordinates, solved on the device like the
        reference's np.linalg.lstsq (minimum norm if rank-deficient).  get_kriging_prediction does
        not go through here: it assembles and solves its systems inside the fused device call."""
        if getattr(self.context, "_model_owner", None) is not self:
            self._push_model()
        cs = np.asarray(coords_spatial, dtype=np.float64).reshape(-1, 2)
        return self.context.kriging_weights(cs[:, 0], cs[:, 1], np.asarray(coords_temporal, dtype=np.float64),
                                            kriging_vector)

    def _get_kriging_weights(self, space, time, min_kriging_points=10, max_kriging_points=100,
                             space_dist_max=None, time_dist_max=None, leave_out_idxs=None):
        if self._kriging_function is None:
            raise ValueError("No Kriging function defined. Did you fit a variogram model?")
        if space_dist_max is None:
            space_dist_max = self._variogram_bins_space[-1]     # stif/predictor.py:787-790
        if time_dist_max is None:
            time_dist_max = self._variogram_bins_time[-1]
        _, _, w, gvec, idx, _ = self._kriging_function(
            space, time, min_kriging_points, max_kriging_points, space_dist_max, time_dist_max,
            leave_out_idxs, True)
        return w, gvec, idx

    def get_kriging_prediction(self, space, time, min_kriging_points=1, max_kriging_points=10,
                               space_dist_max=None, time_dist_max=None, leave_out_idxs=None,
                               batch_size=1000):
        if self._kriging_function is None:
            raise ValueError("No Kriging function defined. Did you fit a variogram model?")
        if space_dist_max is None:
            space_dist_max = self._variogram_bins_space[-1] / 2  # stif/predictor.py:857-860
        if time_dist_max is None:
            time_dist_max = self._variogram_bins_time[-1] / 2
        mean, std = self._kriging_function(
            space, time, min_kriging_points, max_kriging_points, space_dist_max, time_dist_max,
            leave_out_idxs, False)
        return mean, std

    def _get_kriging_prediction_batch(self, space, time, min_kriging_points, max_kriging_points,
                                      space_dist_max, time_dist_max, leave_out_idxs):
        return self._kriging_function(space, time, min_kriging_points, max_kriging_points,
                                      space_dist_max, time_dist_max, leave_out_idxs, False)

    def predict(self, df, kriging_params=dict()):
        covariate_prediction = self.calc_covariate_prediction(df)
        space = df[self._data._space_cols].to_numpy()
        time = df[self._data._time_col].to_numpy()
        kriging_mean, kriging_std = self.get_kriging_prediction(space, time, **kriging_params)
        return covariate_prediction + kriging_mean, kriging_std
