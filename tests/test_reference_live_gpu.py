"""GPU path vs the UNMODIFIED numba reference, live, in the same process on identical inputs
(the north star's wording).  The reference package is the vendored copy `oracle/_ref/`
(oracle/vendor_ref.py, made by __graft_entry__.build() in the build container; git-ignored, shipped
with the snapshot).  Skipped if it or numba is not importable.  The committed golden vectors
(tests/golden/, generated from the same package) cover the same ground without the JIT cost; this
file closes the loop on the box the GPU numbers come from."""
import types
import warnings

import numpy as np
import pandas as pd
import pytest

from oracle import refload
from stif_b200 import Data, Predictor
from stif_b200.utils import get_covariogram, get_variogram

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not refload.available(vendored_only=True):
        pytest.skip("oracle/_ref is absent (run __graft_entry__.build() where /root/reference exists)")
    try:
        import numba  # noqa: F401
        warnings.filterwarnings("ignore")
        return refload.load(vendored_only=True)
    except Exception as exc:       # pragma: no cover
        pytest.skip(f"reference not importable: {exc}")


def test_variogram_live(ref, ctx):
    from stif.utils import get_covariogram as ref_cov, get_variogram as ref_vg
    rng = np.random.default_rng(123)
    n = 4000
    space = rng.uniform(0, 100, (n, 2))
    for t, tmax in ((rng.integers(0, 365, n), 30), (rng.uniform(0, 365, n), 30.0)):
        v = rng.normal(size=n)
        want = ref_vg(space, t, v, 30.0, tmax, 30, 30, None)
        got = get_variogram(space, t, v, 30.0, tmax, 30, 30, None, ctx=ctx)
        assert np.array_equal(got[1], want[1])                                   # counts bit-exact
        assert np.allclose(got[0], want[0], rtol=1e-12, atol=0, equal_nan=True)  # semivariances
        assert got[2:] == want[2:]
    want = ref_cov(space, t, v + 2.0, 25.0, 20.0, 9, 11, None)
    got = get_covariogram(space, t, v + 2.0, 25.0, 20.0, 9, 11, None, ctx=ctx)
    assert np.array_equal(got[1], want[1])
    assert np.allclose(got[0], want[0], rtol=1e-9, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("models,params,k", [
    (("sum_metric", "spherical", "spherical", "spherical"), [400, 60, 500, 0.4, 0.2, 0.3, 0.03, 0.03, 0.03, 5.0], 20),
    (("product_sum", "spherical", "spherical", "spherical"), [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4], 12),
])
def test_kriging_live(ref, ctx, models, params, k):
    rng = np.random.default_rng(321)
    n, n_t = 5000, 150
    df = pd.DataFrame({"x": rng.uniform(0, 150, n), "y": rng.uniform(0, 150, n), "time": rng.uniform(0, 365, n)})
    df["z"] = np.sin(df["x"] / 150) + 0.3 * rng.normal(size=n)
    tsp = rng.uniform(0, 150, (n_t, 2))
    tt = rng.uniform(30, 365, n_t)
    # the reference, parameter vector injected instead of the Nelder-Mead fit
    rp = ref.Predictor(ref.Data(df, space_cols=["x", "y"], time_col="time", predictand_col="z", covariate_cols=[]), None)
    rp._prepare_geostatistics()
    rp._variogram_fit = types.SimpleNamespace(x=np.asarray(params, dtype=float))
    rp._variogram_models = list(models)
    rp._variogram_model_function = rp._create_variogram_model_function()
    rp._kriging_weights_function = rp._create_kriging_weights_function()
    rp._kriging_function = rp._create_kriging_function()
    r_w, r_g, r_idx = rp._get_kriging_weights(tsp, tt, 1, k, 30.0, 30.0, None)
    r_mean, r_std = rp.get_kriging_prediction(tsp, tt, 1, k, 30.0, 30.0, None)
    # this package
    p = Predictor(Data(df, space_cols=["x", "y"], time_col="time", predictand_col="z"), context=ctx)
    p.set_variogram_model(params, *models)
    w, g, idx = p._get_kriging_weights(tsp, tt, 1, k, 30.0, 30.0)
    mean, std = p.get_kriging_prediction(tsp, tt, 1, k, 30.0, 30.0)
    assert np.array_equal(idx, r_idx)                                 # neighbour sets, reference order
    assert np.all(np.abs(g - r_g) <= np.spacing(np.abs(r_g)))         # gamma-vector: 1 float32 ulp
    wscale = np.abs(r_w).max(axis=1, keepdims=True) + 1e-30
    assert np.all(np.abs(w - r_w) <= 2.0 ** -21 * wscale)
    z = df["z"].to_numpy()
    mscale = np.sum(np.abs(r_w.astype(np.float64)) * np.abs(z[r_idx.astype(np.int64)]), axis=1) + 1e-300
    assert np.all(np.abs(mean - r_mean) <= 2.0 ** -21 * mscale)
    assert np.mean(np.abs(mean - r_mean) <= 1e-9 * np.abs(r_mean)) > 0.9
    assert np.array_equal(np.isnan(std), np.isnan(r_std))
    both = ~np.isnan(r_std)
    assert np.allclose(std[both], r_std[both], rtol=2.5e-7, atol=0)
