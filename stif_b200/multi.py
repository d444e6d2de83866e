"""Every GPU of the box from ONE process: `MultiContext` = one libstif_b200 context per device,
each driven by its own host thread (ctypes releases the GIL during the foreign call).

This is SURVEY 8(b)'s `stif_ctx_create(devices, ndev)` on the host side of the C ABI: a plain
`Predictor(...).calc_empirical_variogram()` / `.get_kriging_prediction()` in a notebook uses all
visible GPUs without `torchrun` (`_native.default_context()` builds a MultiContext when more than
one device is visible and the process is not a torch.distributed rank).

* variogram / covariogram: device i accumulates partition i of N of the pair-tile work list
  (`part`/`nparts` of stif_variogram_host); the 16-B-per-bin partial histograms are summed on the
  host in device order -- no collective needed for 14.4 KB.  Counts are integers (bit-exact for
  any N), the sums differ from a one-GPU run by the order of the last N additions only.
* kriging: observations, residuals and the model are replicated, the targets are sharded
  contiguously, the result slices are concatenated (each target's result is independent of the
  sharding: bit-identical to a one-GPU call).

Small calls stay on the first device (launching N uploads costs more than it saves).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _native
from .parallel import shard_bounds

VARIOGRAM_MIN_OBS = 150_000        # below this one GPU finishes in ~2 ms
KRIGING_MIN_TARGETS = 50_000


class MultiContext:
    """Same methods as `_native.Context` for the calls `Predictor` / `Data` / `utils` make."""

    def __init__(self, devices=None):
        if devices is None:
            devices = list(range(max(_native.device_count(), 1)))
        if len(devices) < 1:
            raise ValueError("MultiContext needs at least one device")
        self.devices = [int(d) for d in devices]
        self.contexts = [_native.Context(d) for d in self.devices]
        self._pool = ThreadPoolExecutor(max_workers=len(self.contexts), thread_name_prefix="stif-gpu")
        self._last_kriging = []          # contexts that took part in the last kriging call
        self._last_variogram = []

    # ---- plumbing ----------------------------------------------------------------------
    def __len__(self):
        return len(self.contexts)

    def close(self):
        for c in self.contexts:
            c.close()
        self._pool.shutdown(wait=True)

    def _each(self, fn, contexts=None):
        """fn(i, ctx) on every context concurrently; results in device order; first error raised."""
        ctxs = self.contexts if contexts is None else contexts
        if len(ctxs) == 1:
            return [fn(0, ctxs[0])]
        futs = [self._pool.submit(fn, i, c) for i, c in enumerate(ctxs)]
        return [f.result() for f in futs]

    @property
    def sm_count(self):
        return sum(c.sm_count for c in self.contexts)

    # ---- variogram -----------------------------------------------------------------------
    def variogram_host(self, x, y, t, v, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                       mode=0, mean=0.0, flags=_native.VARIO_PEEL, part=0, nparts=1):
        x, y, t, v = (_native._f64(a) for a in (x, y, t, v))
        ctxs = self.contexts if (len(x) >= VARIOGRAM_MIN_OBS and nparts == 1) else self.contexts[:1]
        self._last_variogram = ctxs
        n = len(ctxs)
        if n == 1:
            return ctxs[0].variogram_host(x, y, t, v, space_dist_max, time_dist_max, n_space_bins,
                                          n_time_bins, mode=mode, mean=mean, flags=flags, part=part, nparts=nparts)
        outs = self._each(lambda i, c: c.variogram_host(
            x, y, t, v, space_dist_max, time_dist_max, n_space_bins, n_time_bins, mode=mode, mean=mean,
            flags=flags, part=i, nparts=n), ctxs)
        counts = outs[0][0].copy()
        sums = outs[0][1].copy()
        for c, s in outs[1:]:              # fixed (device) order
            counts += c
            sums += s
        return counts, sums

    def variogram_bound(self, idx, space_dist_max, time_dist_max, n_space_bins, n_time_bins, mode=0, mean=0.0,
                        flags=_native.VARIO_PEEL, part=0, nparts=1, n_obs=None):
        n_sel = (n_obs or 0) if idx is None else len(idx)
        ctxs = self.contexts if (n_sel >= VARIOGRAM_MIN_OBS and nparts == 1) else self.contexts[:1]
        self._last_variogram = ctxs
        n = len(ctxs)
        if n == 1:
            return ctxs[0].variogram_bound(idx, space_dist_max, time_dist_max, n_space_bins, n_time_bins, mode=mode,
                                           mean=mean, flags=flags, part=part, nparts=nparts)
        outs = self._each(lambda i, c: c.variogram_bound(idx, space_dist_max, time_dist_max, n_space_bins,
                                                         n_time_bins, mode=mode, mean=mean, flags=flags, part=i,
                                                         nparts=n), ctxs)
        counts, sums = outs[0][0].copy(), outs[0][1].copy()
        for c, s in outs[1:]:
            counts += c
            sums += s
        return counts, sums

    def variogram_sampled_host(self, x, y, t, v, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                               n_samples, mode=0, mean=0.0, seed=None, seed_offset=0):
        n_samples = int(n_samples)
        ctxs = self.contexts if n_samples >= 50_000_000 else self.contexts[:1]
        n = len(ctxs)
        if n == 1:
            return ctxs[0].variogram_sampled_host(x, y, t, v, space_dist_max, time_dist_max, n_space_bins,
                                                  n_time_bins, n_samples, mode=mode, mean=mean, seed=seed,
                                                  seed_offset=seed_offset)
        if seed is None:
            import os
            env = os.environ.get("STIF_B200_SEED")
            seed = int(env) if env else int.from_bytes(os.urandom(8), "little")

        def run(i, c):
            b, e = shard_bounds(n_samples, i, n)
            return c.variogram_sampled_host(x, y, t, v, space_dist_max, time_dist_max, n_space_bins, n_time_bins,
                                            e - b, mode=mode, mean=mean, seed=seed, seed_offset=seed_offset * n + i)
        outs = self._each(run, ctxs)
        return sum(o[0] for o in outs), sum(o[1] for o in outs)

    def variogram_stats(self):
        ctxs = self._last_variogram or self.contexts[:1]
        stats = [c.variogram_stats() for c in ctxs]
        return {k: sum(s[k] for s in stats) for k in stats[0]}

    # ---- kriging: replicated state ---------------------------------------------------------
    def kriging_set_obs(self, x, y, t, resid):
        x, y, t, resid = (_native._f64(a) for a in (x, y, t, resid))
        self._each(lambda i, c: c.kriging_set_obs(x, y, t, resid))

    def kriging_set_residuals(self, resid):
        r = _native._f64(resid)
        self._each(lambda i, c: c.kriging_set_residuals(r))

    def kriging_set_model(self, st_model, space_model, time_model, metric_model, params):
        for c in self.contexts:
            c.kriging_set_model(st_model, space_model, time_model, metric_model, params)

    def variogram_model_eval(self, h, t):
        return self.contexts[0].variogram_model_eval(h, t)

    def kriging_weights(self, xs, ys, ts, gamma_vec):
        return self.contexts[0].kriging_weights(xs, ys, ts, gamma_vec)

    def fit_set_grid(self, bins_space, bins_time, empirical, weights):
        return self.contexts[0].fit_set_grid(bins_space, bins_time, empirical, weights)

    def fit_wmse(self, st_model, space_model, time_model, metric_model, params):
        return self.contexts[0].fit_wmse(st_model, space_model, time_model, metric_model, params)

    def kriging_neighbour_pairs(self, tx, ty, tt, space_dist_max, time_dist_max, leave_out=None):
        return self.contexts[0].kriging_neighbour_pairs(tx, ty, tt, space_dist_max, time_dist_max, leave_out)

    # ---- kriging: sharded targets -----------------------------------------------------------
    def kriging_predict_host(self, tx, ty, tt, min_pts, max_pts, space_dist_max, time_dist_max,
                             leave_out=None, flags=_native.KRIG_F32_STORAGE, want_weights=False,
                             out_mean=None, out_std=None):
        tx, ty, tt = (_native._f64(a) for a in (tx, ty, tt))
        n_t = len(tx)
        ctxs = self.contexts if n_t >= KRIGING_MIN_TARGETS else self.contexts[:1]
        self._last_kriging = ctxs
        n = len(ctxs)
        if n == 1:
            return ctxs[0].kriging_predict_host(tx, ty, tt, min_pts, max_pts, space_dist_max, time_dist_max,
                                                leave_out=leave_out, flags=flags, want_weights=want_weights,
                                                out_mean=out_mean, out_std=out_std)
        lo = None if leave_out is None else np.ascontiguousarray(leave_out, dtype=np.int64)
        mean = out_mean if out_mean is not None else np.empty(n_t, dtype=np.float64)
        std = out_std if out_std is not None else np.empty(n_t, dtype=np.float64)

        def run(i, c):
            b, e = shard_bounds(n_t, i, n)
            return c.kriging_predict_host(tx[b:e], ty[b:e], tt[b:e], min_pts, max_pts, space_dist_max,
                                          time_dist_max, leave_out=None if lo is None else lo[b:e], flags=flags,
                                          want_weights=want_weights, out_mean=mean[b:e], out_std=std[b:e])
        outs = self._each(run, ctxs)
        if want_weights:
            return (mean, std) + tuple(np.concatenate([o[q] for o in outs], axis=0) for q in range(2, 6))
        return mean, std

    def kriging_stats(self):
        ctxs = self._last_kriging or self.contexts[:1]
        stats = [c.kriging_stats() for c in ctxs]
        return {k: sum(s[k] for s in stats) for k in stats[0]}

    def kriging_solver_stats(self):
        ctxs = self._last_kriging or self.contexts[:1]
        stats = [c.kriging_solver_stats() for c in ctxs]
        return {"cholesky_to_lu": sum(s["cholesky_to_lu"] for s in stats),
                "cholesky_enabled": all(s["cholesky_enabled"] for s in stats),
                "ldl_enabled": all(s["ldl_enabled"] for s in stats),
                "first_pass_to_lu": sum(s["first_pass_to_lu"] for s in stats)}

    def last_phase_ms(self):
        ctxs = self._last_kriging or self._last_variogram or self.contexts[:1]
        ph = np.array([c.last_phase_ms() for c in ctxs])
        return [float(v) for v in ph.max(axis=0)]
