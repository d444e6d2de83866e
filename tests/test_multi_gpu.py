"""GPU: MultiContext (every visible GPU from one process, stif_b200/multi.py) and the NCCL path of
stif_b200.parallel.  On a one-GPU box the MultiContext tests run over several contexts on the SAME
device (the sharding / merging logic is identical); with more GPUs visible they use them all."""
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from stif_b200 import Data, Predictor, _native, multi

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mctx():
    n = _native.device_count()
    devices = list(range(n)) if n > 1 else [0, 0, 0]
    m = multi.MultiContext(devices)
    yield m
    m.close()


def test_multicontext_variogram_matches_single(ctx, mctx, golden_dir):
    rng = np.random.default_rng(20261019)
    n = 200_000                                     # above multi.VARIOGRAM_MIN_OBS: really sharded
    x, y = rng.uniform(0, 100, n), rng.uniform(0, 100, n)
    t = rng.integers(0, 365, n).astype(np.float64)
    v = rng.normal(size=n)
    c1, s1 = ctx.variogram_host(x, y, t, v, 30.0, 30.0, 30, 30)
    cm, sm = mctx.variogram_host(x, y, t, v, 30.0, 30.0, 30, 30)
    assert len(mctx._last_variogram) == len(mctx) > 1
    assert np.array_equal(c1, cm)                   # bit-exact counts for any number of GPUs
    assert np.all(np.abs(s1 - sm) <= 1e-12 * np.abs(s1))
    assert mctx.variogram_stats()["pairs_in_window"] == ctx.variogram_stats()["pairs_in_window"]
    # the C2 workload against the oracle's whole-workload golden through the multi-GPU path
    g = np.load(os.path.join(golden_dir, "variogram_c2.npz"))
    rng = np.random.default_rng(20261019)           # the C2 generator (bench.py synth_variogram)
    m = 100_000
    x, y = rng.uniform(0, 100, m), rng.uniform(0, 100, m)
    t = rng.integers(0, 365, m).astype(np.float64)
    v = rng.normal(size=m)
    old = multi.VARIOGRAM_MIN_OBS
    multi.VARIOGRAM_MIN_OBS = 1000
    try:
        cm, sm = mctx.variogram_host(x, y, t, v, 30.0, 30.0, 30, 30)
    finally:
        multi.VARIOGRAM_MIN_OBS = old
    assert np.array_equal(cm.astype(np.int64), g["counts"])
    assert np.all(np.abs(sm - g["sums"]) <= 1e-12 * np.abs(g["sums"]))


def test_multicontext_kriging_bit_identical(ctx, mctx):
    rng = np.random.default_rng(3)
    n_obs, n_t = 200_000, 120_000
    x, y, t = rng.uniform(0, 400, n_obs), rng.uniform(0, 400, n_obs), rng.uniform(0, 365, n_obs)
    z = np.sin(x / 150) + 0.3 * rng.normal(size=n_obs)
    tx, ty, tt = rng.uniform(0, 400, n_t), rng.uniform(0, 400, n_t), rng.uniform(30, 365, n_t)
    lo = rng.integers(0, n_obs, n_t)
    model = ("sum_metric", "spherical", "spherical", "spherical")
    params = [400, 60, 500, 0.4, 0.2, 0.3, 0.03, 0.03, 0.03, 5.0]
    for c in (ctx, mctx):
        c.kriging_set_obs(x, y, t, z)
        c.kriging_set_model(*model, params)
    a = ctx.kriging_predict_host(tx, ty, tt, 1, 20, 30.0, 30.0, leave_out=lo, want_weights=True)
    b = mctx.kriging_predict_host(tx, ty, tt, 1, 20, 30.0, 30.0, leave_out=lo, want_weights=True)
    assert len(mctx._last_kriging) == len(mctx)
    for q in range(6):
        assert np.array_equal(a[q], b[q], equal_nan=True), q
    assert ctx.kriging_stats() == mctx.kriging_stats()
    # residual refresh reaches every device
    z2 = z + 1.0
    ctx.kriging_set_residuals(z2)
    mctx.kriging_set_residuals(z2)
    a = ctx.kriging_predict_host(tx, ty, tt, 1, 20, 30.0, 30.0)
    b = mctx.kriging_predict_host(tx, ty, tt, 1, 20, 30.0, 30.0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_predictor_over_multicontext(mctx, ctx):
    """The reference-facing API on top of a MultiContext == on top of one context."""
    rng = np.random.default_rng(11)
    n = 160_000
    df = pd.DataFrame({"x": rng.uniform(0, 300, n), "y": rng.uniform(0, 300, n),
                       "time": rng.integers(0, 365, n), "z": rng.normal(size=n)})
    res = []
    for c in (ctx, mctx):
        p = Predictor(Data(df, space_cols=["x", "y"], time_col="time", predictand_col="z"), context=c)
        p.calc_empirical_variogram(space_dist_max=30.0, time_dist_max=30, n_space_bins=12, n_time_bins=10)
        p.set_variogram_model([400, 60, 500, 0.4, 0.2, 0.3, 0.03, 0.03, 0.03, 5.0])
        tsp = rng.uniform(0, 300, (60_000, 2)) if not res else tsp
        tt = rng.uniform(30, 365, 60_000) if not res else tt
        res.append((p._variogram, p._variogram_samples_per_bin, *p.get_kriging_prediction(tsp, tt, 1, 16, 30.0, 30.0)))
    assert np.array_equal(res[0][1], res[1][1])
    assert np.allclose(res[0][0], res[1][0], rtol=1e-12, atol=0, equal_nan=True)
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3], equal_nan=True)


def test_default_context_uses_all_visible_gpus():
    """No torchrun, no arguments: default_context() spans every visible GPU; CUDA_VISIBLE_DEVICES
    and STIF_B200_DEVICES narrow the set; a torch.distributed rank gets exactly one."""
    code = ("import os,sys; sys.path.insert(0, %r); from stif_b200 import _native; c=_native.default_context(); "
            "print(type(c).__name__, getattr(c, 'devices', None), _native.device_count())" % ROOT)

    def run(env):
        e = dict(os.environ)
        for k in ("LOCAL_RANK", "RANK", "WORLD_SIZE", "STIF_B200_DEVICE", "STIF_B200_DEVICES"):
            e.pop(k, None)
        e.update(env)
        return subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=300).stdout.split()

    n = _native.device_count()
    out = run({})
    assert out[0] == ("MultiContext" if n > 1 else "Context")
    assert run({"CUDA_VISIBLE_DEVICES": "0"})[0] == "Context"
    assert run({"STIF_B200_DEVICES": "0,0"})[0] == "MultiContext"        # explicit list (same device twice)
    assert run({"WORLD_SIZE": "2", "LOCAL_RANK": "0"})[0] == "Context"
