"""`Data`: the observation table -- host side, same constructor and accessors as
stif/data.py:45-168 of the reference (column selection, min-max normalisation, covariate
transforms).  The reference's neighbour pre-filter `Data.get_kriging_idxs`
(stif/data.py:170-212) is NOT here: the dense n_obs x batch filter is replaced by the
on-device space-time grid search in csrc/kriging.cu.
"""
import math
from functools import cached_property

import numpy as np
import pandas as pd


def sinusoidal_feature_transform(x, n_freqs=6, full_circle=None):
    """sin/cos features of a 1-D coordinate (stif/data.py:10-42): column i is
    sin (i even) or cos (i odd) of (1 + i//2) * 2*pi*x/full_circle."""
    x = np.asarray(x, dtype=float)
    if full_circle is None:
        full_circle = x.max() - x.min()
    phase = x / full_circle * 2 * math.pi
    out = np.empty((len(x), n_freqs), dtype=float)
    for i in range(n_freqs):
        harmonic = 1 + i // 2
        out[:, i] = np.sin(phase * harmonic) if i % 2 == 0 else np.cos(phase * harmonic)
    return out


class Data:
    """Unstructured space-time observations, one row per observation."""

    def __init__(self, df: pd.DataFrame, space_cols=["longitude", "latitude"], time_col=None,
                 predictand_col="predictand", covariate_cols=list(), normalize=True,
                 covariate_transformations=dict()):
        self._training_df = df
        self._space_cols = space_cols
        self._time_col = time_col
        self._predictand_col = predictand_col
        self._covariate_cols = covariate_cols
        self._normalize = normalize
        self._covariate_transformations = covariate_transformations

    @cached_property
    def _normalization_bounds(self):
        bounds = {}
        for col in self._training_df.columns:
            series = self._training_df[col]
            if series.dtype.kind in "iuf":
                bounds[col] = (series.min(), series.max())
        return bounds

    def normalize(self, array, col_name):
        lo, hi = self._normalization_bounds[col_name]
        return (array - lo) / (hi - lo)

    @property
    def space_coords(self):
        """(n, 2) spatial coordinates."""
        return self._training_df[self._space_cols].to_numpy()

    @property
    def time_coords(self):
        """(n,) temporal coordinates (int64 or float64)."""
        return self._training_df[self._time_col].to_numpy()

    @property
    def predictand(self):
        """(n,) predictand (float or bool)."""
        return self._training_df[self._predictand_col].to_numpy()

    def prepare_covariates(self, df):
        """Covariate matrix of `df`: per covariate column its transformed features (if a
        transformation is registered) followed by the (optionally normalised) column."""
        blocks = [np.empty((len(df), 0), dtype=float)]
        for col in self._covariate_cols:
            values = df[col].to_numpy(copy=True)
            if col in self._covariate_transformations:
                feats = np.asarray(self._covariate_transformations[col](values))
                blocks.append(feats.reshape(len(df), -1))
            if self._normalize:
                values = self.normalize(values, col)
            blocks.append(np.asarray(values, dtype=float).reshape(len(df), 1))
        return np.concatenate(blocks, axis=1)

    def get_training_covariates(self):
        return self.prepare_covariates(self._training_df)

    def get_kriging_idxs(self, space, time, space_dist_max, time_dist_max, leave_out_idxs, context=None):
        """Indices of the training data within the spatial and temporal distance of each target
        (stif/data.py:170-212): (P, 2) int64 [obs_idx, target_idx], ordered like np.nonzero.
        Found on the GPU through the space-time grid instead of the dense n_obs x n_targets
        distance matrices.  (The Predictor does not call this: its kriging path fuses the
        search with the selection of the k lowest-gamma neighbours.)"""
        from . import _native
        ctx = context if context is not None else _native.default_context()
        if getattr(ctx, "_obs_owner", None) is not self:
            x, y = _native.split_space(self.space_coords)
            t = self.time_coords
            ctx.kriging_set_obs(x, y, t, np.zeros(len(t)))
            ctx._obs_owner = self
        tx, ty = _native.split_space(np.asarray(space).reshape(-1, 2))
        return ctx.kriging_neighbour_pairs(tx, ty, np.asarray(time).reshape(-1), space_dist_max,
                                           time_dist_max, leave_out_idxs)
