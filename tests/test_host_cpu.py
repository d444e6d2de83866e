"""CPU: host-side logic -- variogram-model functions, Data, Predictor call-order errors,
and the multi-rank layer (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

from oracle import kriging as ok
from oracle import variogram as ov
from stif_b200 import Data, Predictor, parallel, sinusoidal_feature_transform
from stif_b200 import variogram_models as vm
from stif_b200.utils import calc_distance_matrix_1d, calc_distance_matrix_2d


def test_distance_matrices():
    # reference tests/test_utils.py:6-29
    vec = np.array([1, 2, 3, 4, 5])
    want = np.array([[0, 1, 2, 3, 4], [1, 0, 1, 2, 3], [2, 1, 0, 1, 2], [3, 2, 1, 0, 1], [4, 3, 2, 1, 0]])
    assert np.array_equal(calc_distance_matrix_1d(vec), want)
    pts = np.random.default_rng(0).normal(size=(6, 2))
    d = calc_distance_matrix_2d(pts)
    for i in range(6):
        for j in range(6):
            assert np.isclose(d[i, j], np.linalg.norm(pts[i] - pts[j]))


@pytest.mark.parametrize("st,n", [("sum", 6), ("product", 6), ("product_sum", 7), ("metric", 4), ("sum_metric", 10)])
@pytest.mark.parametrize("marg", ["spherical", "gaussian"])
def test_host_models_match_oracle(st, n, marg):
    rng = np.random.default_rng(3)
    x = rng.uniform(0.1, 50, n)
    h = rng.uniform(0, 80, 200)
    t = rng.uniform(0, 80, 200)
    got = vm.variogram_model_dict[st](h, t, x, vm.variogram_model_dict[marg], vm.variogram_model_dict[marg],
                                      vm.variogram_model_dict[marg])
    want = ok.gamma((st, marg, marg, marg), x, h, t)
    assert np.allclose(got, want, rtol=1e-13, atol=0)


def test_fit_helpers():
    rng = np.random.default_rng(0)
    ev = rng.uniform(1, 2, (5, 4))
    for st, n in vm.N_PARAMS.items():
        assert len(vm.get_initial_parameters(st, ev, 10.0, 5.0, 0.3)) == n
    w = vm.calc_weights(np.arange(1, 6.0), np.arange(1, 5.0), 0.5, np.full((5, 4), 7))
    assert w.shape == (5, 4) and w.max() == 1.0
    x = vm.get_initial_parameters("sum_metric", ev, 10.0, 5.0, 0.3)
    f = vm.weighted_mean_square_error(x, vm.sum_metric_model, vm.spherical, vm.spherical, vm.spherical,
                                      np.arange(1, 6.0), np.arange(1, 5.0), ev, w)
    assert np.isfinite(f)


def _df(n=50):
    rng = np.random.default_rng(1)
    return pd.DataFrame({"x": rng.uniform(0, 1, n), "y": rng.uniform(0, 1, n),
                         "time": np.arange(n), "z": rng.normal(size=n)})


def test_data_accessors():
    df = _df()
    d = Data(df, space_cols=["x", "y"], time_col="time", predictand_col="z", covariate_cols=["x", "time"],
             covariate_transformations={"time": lambda c: sinusoidal_feature_transform(c, n_freqs=4)})
    assert d.space_coords.shape == (50, 2) and d.time_coords.dtype == np.int64
    X = d.get_training_covariates()
    assert X.shape == (50, 1 + 4 + 1)
    assert X[:, 0].min() == 0.0 and X[:, 0].max() == 1.0      # normalised
    f = sinusoidal_feature_transform(np.linspace(0, 1, 11), n_freqs=3, full_circle=1.0)
    assert np.allclose(f[:, 0], np.sin(2 * np.pi * np.linspace(0, 1, 11)))
    assert np.allclose(f[:, 1], np.cos(2 * np.pi * np.linspace(0, 1, 11)))


def test_predictor_call_order_errors():
    df = _df()
    p = Predictor(Data(df, space_cols=["x", "y"], time_col="time", predictand_col="z"))
    with pytest.raises(ValueError):
        p.get_kriging_prediction(np.zeros((1, 2)), np.zeros(1))     # predictor.py:852-854
    with pytest.raises(ValueError):
        p.save_empirical_variogram("/tmp/x.npz")                     # predictor.py:510-511
    with pytest.raises(ValueError):
        p.get_cross_val_metric(lambda a, b: 0)
    with pytest.raises(ValueError):
        p._get_variogram_model_grid()
    assert np.array_equal(p.get_residuals(), df["z"].to_numpy())     # no covariate model


def test_shard_bounds():
    for n in (0, 1, 7, 100):
        for ws in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, ws) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        n = 600
        space = rng.uniform(0, 10, (n, 2))
        t = rng.integers(0, 30, n)
        v = rng.normal(size=n)
        # each rank bins its share of the rows (stand-in for the kernel's partition), then
        # the histograms meet in one all-reduce
        b, e = parallel.shard_bounds(n, rank, world)
        h, c = ov.raw(space, t, v, 3.0, 10, 6, 5, i_begin=b, i_end=e)
        counts, sums = parallel.allreduce_histograms(c.astype(np.uint64), h)
        # the same through the public multi-rank entry point (non-NCCL branch: the host call per rank
        # with part / nparts, then the ONE fused float64 all-reduce); the context is a stand-in whose
        # "kernel" is the oracle on this rank's row block
        class _Ctx:
            def variogram_host(self, x, y, t_, v_, smax, tmax, ns, nt, mode=0, mean=0.0, flags=1, part=0, nparts=1):
                assert (part, nparts) == (rank, world)
                bb, ee = parallel.shard_bounds(len(x), part, nparts)
                hh, cc = ov.raw(np.column_stack([x, y]), t_, v_, smax, tmax, ns, nt, i_begin=bb, i_end=ee)
                return cc.astype(np.uint64), hh
        counts2, sums2 = parallel.variogram_all_ranks(_Ctx(), space[:, 0].copy(), space[:, 1].copy(),
                                                      t.astype(np.float64), v, 3.0, 10, 6, 5, 0, 0.0)
        assert np.array_equal(counts2, counts) and np.array_equal(sums2, sums)
        assert counts2.dtype == np.uint64 and counts2.shape == (6, 5)
        # kriging: shard targets, gather
        vals = np.arange(101, dtype=np.float64)
        idx = np.arange(101 * 3, dtype=np.uint32).reshape(101, 3)
        got = parallel.kriging_all_ranks(lambda lo, hi: (vals[lo:hi] * 2, idx[lo:hi]), 101)
        q.put((rank, counts, sums, got[0], got[1]))
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    n = 600
    space = rng.uniform(0, 10, (n, 2))
    t = rng.integers(0, 30, n)
    v = rng.normal(size=n)
    h, c = ov.raw(space, t, v, 3.0, 10, 6, 5)
    for rank, counts, sums, m, idx in results:
        assert np.array_equal(counts.astype(np.float64), c)       # bit-exact counts
        assert np.allclose(sums, h, rtol=1e-13)
        assert np.array_equal(m, np.arange(101) * 2.0)
        assert np.array_equal(idx, np.arange(303, dtype=np.uint32).reshape(101, 3))


class _RecordingContext:
    """Stand-in for _native.Context that records what reaches the device (no GPU needed)."""

    def __init__(self):
        self.resid = None
        self.calls = []

    def kriging_set_obs(self, x, y, t, resid):
        self.calls.append("set_obs")
        self.resid = np.array(resid, copy=True)

    def kriging_set_residuals(self, resid):
        self.calls.append("set_residuals")
        self.resid = np.array(resid, copy=True)


def test_residual_upload_follows_every_refresh():
    """the device residual cache was keyed on id(self._residuals); CPython reuses the
    address of a freed array, so two refreshes between kriging calls could leave stale residuals
    on the GPU.  The key is now a version counter bumped on every assignment."""
    from sklearn.linear_model import LinearRegression
    df = _df(200)
    df["z"] = 3.0 * df["x"] + np.random.default_rng(2).normal(size=200)
    data = Data(df, space_cols=["x", "y"], time_col="time", predictand_col="z", covariate_cols=["x"])
    ctx = _RecordingContext()
    p = Predictor(data, LinearRegression(), context=ctx)
    p.fit_covariate_model(np.arange(0, 100))
    p._prepare_geostatistics()
    p._upload_observations()
    assert ctx.calls == ["set_obs"] and np.array_equal(ctx.resid, p._residuals)
    p._upload_observations()                       # nothing changed: no transfer
    assert ctx.calls == ["set_obs"]
    for rows in (np.arange(100, 200), np.arange(50, 150), np.arange(0, 200)):
        # refresh TWICE between predictions (predict -> refit -> variogram -> covariogram -> predict):
        # the first refreshed array is freed, the second may land on its address
        p.fit_covariate_model(rows)
        p._prepare_geostatistics()
        p._prepare_geostatistics()
        p._upload_observations()
        assert ctx.calls[-1] == "set_residuals"
        assert np.array_equal(ctx.resid, p._residuals)
        assert np.array_equal(ctx.resid, p.get_residuals())
    n_calls = len(ctx.calls)
    p._residuals = p._residuals.copy()             # direct assignment counts as a refresh too
    p._upload_observations()
    assert len(ctx.calls) == n_calls + 1
