"""GPU parity: libstif_b200 kriging (neighbour search, top-k, assemble + solve, epilogue)
vs the CPU oracle, through the C ABI (stif_kriging_predict_host).

Tolerances (north star): neighbour sets identical (ties broken by index); kriging mean and
variance within 1e-9 relative in float64 mode.  In float32-storage mode (the stock
reference's arithmetic) weights/gammas must agree to 1 float32 ulp and mean/std to the
float32 rounding that implies."""
import numpy as np
import pytest

from oracle import kriging as ok
from stif_b200 import _native

pytestmark = pytest.mark.gpu

SUM_METRIC = ("sum_metric", "spherical", "spherical", "spherical")
PRODUCT_SUM = ("product_sum", "spherical", "spherical", "spherical")
P_SUM_METRIC = [400, 60, 500, 0.4, 0.2, 0.3, 0.03, 0.03, 0.03, 5.0]
P_PRODUCT_SUM = [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4]
RTOL = 1e-9


def synth(n_obs, n_tgt, seed=20261020, extent=300.0):
    rng = np.random.default_rng(seed)
    space = rng.uniform(0, extent, (n_obs, 2))
    t = rng.uniform(0, 365, n_obs)
    z = np.sin(space[:, 0] / 150) + np.cos(space[:, 1] / 200) + 0.5 * np.sin(t / 30) + 0.3 * rng.normal(size=n_obs)
    tsp = rng.uniform(0, extent, (n_tgt, 2))
    tt = rng.uniform(30, 365, n_tgt)
    return space, t, z, tsp, tt


SOLVER_FLAGS = {"auto": 0, "lu": _native.KRIG_FORCE_LU, "chol": _native.KRIG_TRY_CHOL,
                "ldl": _native.KRIG_TRY_LDL}


def gpu_predict(ctx, space, t, z, tsp, tt, models, params, min_pts, k, smax, tmax, leave=None, f32=True,
                solver="auto"):
    x, y = _native.split_space(space)
    ctx.kriging_set_obs(x, y, t, z)
    ctx.kriging_set_model(*models, params)
    tx, ty = _native.split_space(tsp)
    flags = (_native.KRIG_F32_STORAGE if f32 else 0) | SOLVER_FLAGS[solver]
    mean, std, w, g, idx, nused = ctx.kriging_predict_host(tx, ty, tt, min_pts, k, smax, tmax, leave_out=leave,
                                                           flags=flags, want_weights=True)
    return dict(mean=mean, std=std, w=w, gamma=g, idx=idx, n_used=nused)


def compare(g, o, z, f32):
    n = len(o["mean"])
    ok_rows = o["tie"] == 0
    assert ok_rows.mean() > 0.9
    assert np.array_equal(g["n_used"][ok_rows], o["n_used"][ok_rows])
    # neighbour sets (and their order) identical
    assert np.array_equal(g["idx"][ok_rows], o["idx"][ok_rows])
    # gamma-vector: same formula, float32-rounded -> allow 1 ulp
    assert np.all(np.abs(g["gamma"][ok_rows] - o["gamma"][ok_rows]) <= np.spacing(np.abs(o["gamma"][ok_rows])))
    wscale = np.abs(o["w"]).max(axis=1, keepdims=True) + 1e-30
    solved = o["n_used"] > 0
    nonsing = ok_rows & solved & (o["rank"] == o["n_used"] + 1)
    assert nonsing.sum() > 0.9 * solved.sum()
    skipped = ~solved
    assert np.all(g["mean"][skipped] == 0) and np.all(g["std"][skipped] == 0)
    # weights agree to float32 resolution of the row
    assert np.all(np.abs(g["w"][nonsing] - o["w"][nonsing]) <= 2.0 ** -22 * wscale[nonsing])
    resid_abs = np.abs(z[o["idx"].astype(np.int64)])
    mscale = np.sum(np.abs(o["w"].astype(np.float64)) * resid_abs, axis=1) + 1e-300
    vscale = np.sum(np.abs(o["w"].astype(np.float64) * o["gamma"]), axis=1) + 1e-300
    dm = np.abs(g["mean"] - o["mean"])[nonsing]
    var_g, var_o = g["std"] ** 2, o["std"] ** 2
    both_nan = np.isnan(var_g) & np.isnan(var_o)
    dv = np.where(both_nan, 0.0, np.abs(var_g - var_o))[nonsing]
    if f32:
        # a weight may land on the other side of a float32 rounding boundary
        assert np.all(dm <= 2.0 ** -22 * mscale[nonsing])
        assert np.mean(dm <= RTOL * mscale[nonsing]) > 0.8
        assert np.all(dv <= 2.0 ** -20 * vscale[nonsing])
    else:
        assert np.all(dm <= RTOL * mscale[nonsing]), dm.max()
        assert np.all(dv <= RTOL * vscale[nonsing]), dv.max()


@pytest.mark.parametrize("solver", ["auto", "lu", "chol", "ldl"])
@pytest.mark.parametrize("f32", [False, True])
@pytest.mark.parametrize("models,params,k", [
    (SUM_METRIC, P_SUM_METRIC, 50),
    (PRODUCT_SUM, P_PRODUCT_SUM, 20),
    (("sum_metric", "gaussian", "spherical", "gaussian"), P_SUM_METRIC, 10),
    (("sum", "spherical", "gaussian", "spherical"), [300, 60, 0.8, 0.7, 0.15, 0.10], 16),
    (("metric", "spherical", "spherical", "spherical"), [500, 1.0, 0.05, 5.0], 33),
    (("product", "spherical", "spherical", "spherical"), [300, 60, 0.8, 0.7, 0.15, 0.10], 8),
])
def test_models(ctx, models, params, k, f32, solver):
    space, t, z, tsp, tt = synth(20000, 400)
    g = gpu_predict(ctx, space, t, z, tsp, tt, models, params, 1, k, 30.0, 30.0, f32=f32, solver=solver)
    o = ok.predict(space, t, z, tsp, tt, models, params, 1, k, 30.0, 30.0, f32_storage=f32)
    compare(g, o, z, f32)
    ss = ctx.kriging_solver_stats()
    valid = models[0] in ("sum", "metric", "sum_metric")
    if solver == "lu":
        assert not ss["cholesky_enabled"] and not ss["ldl_enabled"]
    elif solver == "ldl":
        # every model through the symmetric indefinite solver; distinct points: nothing handed on
        assert ss["ldl_enabled"] and not ss["cholesky_enabled"] and ss["first_pass_to_lu"] == 0
    elif solver == "chol" or valid:
        assert ss["cholesky_enabled"] and not ss["ldl_enabled"]
        if models[0] in ("metric", "sum_metric"):
            # valid variogram, distinct points: positive definite, nothing handed to the LU
            assert ss["cholesky_to_lu"] == 0
    else:
        # product / product_sum: not a valid variogram -> Bunch-Kaufman LDL' of the bordered system
        assert ss["ldl_enabled"] and not ss["cholesky_enabled"] and ss["first_pass_to_lu"] == 0


def test_cholesky_hands_indefinite_systems_to_lu(ctx):
    """product_sum is not a valid variogram: forcing the covariance-form kernel must detect the
    non-positive pivots and give the same answers through the LU."""
    space, t, z, tsp, tt = synth(20000, 300, seed=21)
    g1 = gpu_predict(ctx, space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, 24, 30.0, 30.0, f32=False,
                     solver="chol")
    ss = ctx.kriging_solver_stats()
    assert ss["cholesky_enabled"] and ss["cholesky_to_lu"] > 0
    g2 = gpu_predict(ctx, space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, 24, 30.0, 30.0, f32=False,
                     solver="lu")
    o = ok.predict(space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, 24, 30.0, 30.0, f32_storage=False)
    compare(g1, o, z, False)
    compare(g2, o, z, False)


@pytest.mark.parametrize("models,params,k,radii", [
    # every Cholesky kernel variant: {sum, metric, sum_metric} x spherical, with and without the
    # range clamp (radii beyond the ranges), two- and four-column-slot kernels (k <= 64 / > 64)
    (("sum", "spherical", "spherical", "spherical"), [300, 60, 0.8, 0.7, 0.15, 0.10], 40, (30.0, 30.0)),
    (("sum", "spherical", "spherical", "spherical"), [20, 10, 0.8, 0.7, 0.15, 0.10], 40, (30.0, 30.0)),
    (("metric", "spherical", "spherical", "spherical"), [100, 1.0, 0.05, 5.0], 24, (30.0, 30.0)),
    (SUM_METRIC, [40, 20, 50, 0.4, 0.2, 0.3, 0.03, 0.03, 0.03, 5.0], 50, (30.0, 30.0)),
    (SUM_METRIC, [40, 20, 50, 0.4, 0.2, 0.3, 0.03, 0.03, 0.03, 5.0], 70, (40.0, 40.0)),
    (("sum", "spherical", "spherical", "spherical"), [300, 60, 0.8, 0.7, 0.15, 0.10], 90, (40.0, 40.0)),
    (("sum_metric", "gaussian", "gaussian", "spherical"), P_SUM_METRIC, 80, (40.0, 40.0)),
    # marginal nuggets that cancel (what unconstrained Nelder-Mead fits return): a valid model
    (SUM_METRIC, [400, 60, 500, 0.4, 0.2, 0.3, -1000.0, -500.0, 1500.09, 5.0], 50, (30.0, 30.0)),
    (("sum", "spherical", "spherical", "spherical"), [300, 60, 0.8, 0.7, -70.0, 70.25], 70, (40.0, 40.0)),
    (("sum_metric", "spherical", "gaussian", "spherical"), [400, 60, 500, 0.4, 0.2, 0.3, -10.0, -5.0, 15.09, 5.0], 30,
     (30.0, 30.0)),
])
def test_cholesky_variants(ctx, models, params, k, radii):
    space, t, z, tsp, tt = synth(30000, 200, seed=31, extent=200.0)
    g = gpu_predict(ctx, space, t, z, tsp, tt, models, params, 1, k, radii[0], radii[1], f32=False)
    ss = ctx.kriging_solver_stats()
    assert ss["cholesky_enabled"]
    o = ok.predict(space, t, z, tsp, tt, models, params, 1, k, radii[0], radii[1], f32_storage=False)
    compare(g, o, z, False)
    # and the same through the pivoted LU
    g2 = gpu_predict(ctx, space, t, z, tsp, tt, models, params, 1, k, radii[0], radii[1], f32=False, solver="lu")
    compare(g2, o, z, False)


@pytest.mark.parametrize("models,params,k,radii", [
    # pivoted-LU assembly variants: polynomial product / product_sum (inside the ranges) and the
    # verbatim formula (beyond them, or other marginals)
    (PRODUCT_SUM, P_PRODUCT_SUM, 50, (30.0, 30.0)),
    (PRODUCT_SUM, [40, 20, 0.8, 0.7, 0.15, 0.10, 0.4], 50, (30.0, 30.0)),
    (("product", "spherical", "spherical", "spherical"), [300, 60, 0.8, 0.7, 0.15, 0.10], 70, (40.0, 40.0)),
    (("product_sum", "gaussian", "spherical", "spherical"), P_PRODUCT_SUM, 33, (30.0, 30.0)),
])
def test_lu_variants(ctx, models, params, k, radii):
    space, t, z, tsp, tt = synth(30000, 200, seed=33, extent=200.0)
    g = gpu_predict(ctx, space, t, z, tsp, tt, models, params, 1, k, radii[0], radii[1], f32=False)
    ss = ctx.kriging_solver_stats()
    assert not ss["cholesky_enabled"] and ss["ldl_enabled"] and ss["first_pass_to_lu"] == 0
    o = ok.predict(space, t, z, tsp, tt, models, params, 1, k, radii[0], radii[1], f32_storage=False)
    compare(g, o, z, False)
    # and the same through the pivoted LU (same assembly expressions, general elimination)
    g2 = gpu_predict(ctx, space, t, z, tsp, tt, models, params, 1, k, radii[0], radii[1], f32=False, solver="lu")
    assert not ctx.kriging_solver_stats()["ldl_enabled"]
    compare(g2, o, z, False)
    # the two solvers factor the same matrix: weights agree far below the float32 storage step
    wscale = np.abs(g2["w"]).max(axis=1, keepdims=True) + 1e-30
    assert np.all(np.abs(g["w"] - g2["w"]) <= 2.0 ** -22 * wscale)


@pytest.mark.parametrize("k", [1, 2, 3, 7, 30, 31, 62, 63, 94, 95, 100, 126, 127, 128])
def test_ldl_row_group_boundaries(ctx, k):
    """The LDL' kernel is instantiated per number of 32-row groups of the k + 2 rows: both sides of
    every boundary, for the indefinite model that takes it by default and for a valid one."""
    space, t, z, tsp, tt = synth(40000, 96, seed=40 + k, extent=300.0)
    for models, params in ((PRODUCT_SUM, P_PRODUCT_SUM), (SUM_METRIC, P_SUM_METRIC)):
        g = gpu_predict(ctx, space, t, z, tsp, tt, models, params, 1, k, 45.0, 45.0, f32=False, solver="ldl")
        ss = ctx.kriging_solver_stats()
        assert ss["ldl_enabled"] and ss["first_pass_to_lu"] == 0
        o = ok.predict(space, t, z, tsp, tt, models, params, 1, k, 45.0, 45.0, f32_storage=False)
        if k >= 100:
            assert (o["n_used"] == k).mean() > 0.5
        compare(g, o, z, False)


def test_ldl_hands_duplicates_to_lu_and_minnorm(ctx):
    """Duplicate observations make the bordered system exactly singular: the LDL' kernel must see the
    vanishing pivot, hand the target to the LU, which flags it for the minimum-norm kernel; the
    results equal those of the LU-only path bit for bit (same second and third pass)."""
    space, t, z, tsp, tt = synth(6000, 150, seed=77, extent=120.0)
    space = np.concatenate([space, space[:3000]])
    t = np.concatenate([t, t[:3000]])
    z = np.concatenate([z, z[:3000]])
    a = gpu_predict(ctx, space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, 24, 30.0, 30.0, f32=False)
    ss, st = ctx.kriging_solver_stats(), ctx.kriging_stats()
    assert ss["ldl_enabled"] and ss["first_pass_to_lu"] > 0 and st["targets_singular"] > 0
    assert ss["first_pass_to_lu"] >= st["targets_singular"]
    b = gpu_predict(ctx, space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, 24, 30.0, 30.0, f32=False,
                    solver="lu")
    st2 = ctx.kriging_stats()
    assert st2["targets_singular"] == st["targets_singular"]
    # targets the LDL' kernel kept: same matrix, another elimination order
    wscale = np.abs(b["w"]).max(axis=1, keepdims=True) + 1e-30
    assert np.all(np.abs(a["w"] - b["w"]) <= 2.0 ** -20 * wscale)
    assert np.array_equal(a["idx"], b["idx"])
    assert np.allclose(a["mean"], b["mean"], rtol=0, atol=1e-6 * np.abs(z).max())


def test_k100_many_candidates(ctx):
    # > 512 candidates per target exercises the streaming select
    space, t, z, tsp, tt = synth(60000, 120, seed=7)
    g = gpu_predict(ctx, space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, 100, 60.0, 60.0, f32=False)
    o = ok.predict(space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, 100, 60.0, 60.0, f32_storage=False)
    assert o["n_cand"].max() > 600
    compare(g, o, z, False)


def test_min_points_and_empty(ctx):
    space, t, z, tsp, tt = synth(3000, 300, seed=3)
    tt[:50] = -5.0            # before every observation: causal filter leaves nothing
    tsp[50:60] = 1e6          # far outside the grid
    g = gpu_predict(ctx, space, t, z, tsp, tt, SUM_METRIC, P_SUM_METRIC, 10, 12, 30.0, 30.0)
    o = ok.predict(space, t, z, tsp, tt, SUM_METRIC, P_SUM_METRIC, 10, 12, 30.0, 30.0)
    assert np.all(g["n_used"][:60] == -1) and np.all(g["mean"][:60] == 0) and np.all(g["std"][:60] == 0)
    assert (o["n_used"] == -1).sum() > 60
    compare(g, o, z, True)


@pytest.mark.parametrize("k", [12, 100])
def test_ldl_ragged_and_empty_targets(ctx, k):
    """The LDL' kernel (one warp per target at k = 12, four at k = 100) on targets without any
    candidate, below min_kriging_points, and with anything between 10 and k neighbours."""
    space, t, z, tsp, tt = synth(3000, 300, seed=3)
    tt[:50] = -5.0            # before every observation: causal filter leaves nothing
    tsp[50:60] = 1e6          # far outside the grid
    g = gpu_predict(ctx, space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 10, k, 60.0, 60.0, f32=False)
    ss = ctx.kriging_solver_stats()
    assert ss["ldl_enabled"] and ss["first_pass_to_lu"] == 0
    o = ok.predict(space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 10, k, 60.0, 60.0, f32_storage=False)
    assert np.all(g["n_used"][:60] == -1) and np.all(g["mean"][:60] == 0) and np.all(g["std"][:60] == 0)
    used = o["n_used"][o["n_used"] > 0]
    assert used.min() < min(k, 30) and used.max() == min(k, used.max())   # ragged systems (and full ones at k = 12)
    assert k > 12 or used.max() == k
    compare(g, o, z, False)


def test_ldl_fewer_observations_than_k(ctx):
    space, t, z, tsp, tt = synth(12, 40, seed=8, extent=20.0)
    tt[:] = 400.0
    for k in (50, 100):
        g = gpu_predict(ctx, space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, k, 1e3, 1e3, f32=False)
        assert ctx.kriging_solver_stats()["ldl_enabled"]
        o = ok.predict(space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, k, 1e3, 1e3, f32_storage=False)
        assert np.all(o["n_used"] == 12)
        compare(g, o, z, False)


def test_leave_out(ctx):
    space, t, z, _, _ = synth(5000, 1, seed=5)
    sel = np.arange(0, 5000, 10)
    tsp, tt = space[sel], t[sel]
    g = gpu_predict(ctx, space, t, z, tsp, tt, SUM_METRIC, P_SUM_METRIC, 1, 20, 60.0, 60.0, leave=sel, f32=False)
    o = ok.predict(space, t, z, tsp, tt, SUM_METRIC, P_SUM_METRIC, 1, 20, 60.0, 60.0, leave_out_idxs=sel,
                   f32_storage=False)
    for i, s in enumerate(sel):
        assert s not in g["idx"][i][:max(g["n_used"][i], 0)]
    compare(g, o, z, False)
    # without leave-out the observation itself (h=0, t=0) is the first neighbour
    g2 = gpu_predict(ctx, space, t, z, tsp, tt, SUM_METRIC, P_SUM_METRIC, 1, 20, 60.0, 60.0, f32=False)
    assert np.mean(np.any(g2["idx"] == sel[:, None], axis=1)) > 0.99


def test_integer_times_strict_radii(ctx):
    # lattice in space and integer days: candidates exactly ON the radii must be excluded
    gx, gy = np.meshgrid(np.arange(12.0), np.arange(12.0))
    pts = np.column_stack([gx.ravel(), gy.ravel()])
    days = np.arange(40)
    space = np.repeat(pts, len(days), axis=0)
    t = np.tile(days, len(pts)).astype(np.int64)
    rng = np.random.default_rng(2)
    z = rng.normal(size=len(t))
    tsp = pts[rng.integers(0, len(pts), 200)] + 0.0
    tt = rng.integers(5, 40, 200)
    for f32 in (True, False):
        g = gpu_predict(ctx, space, t, z, tsp, tt, PRODUCT_SUM, [6, 9, 0.8, 0.7, 0.15, 0.10, 0.4], 1, 30, 3.0, 4,
                        f32=f32)
        o = ok.predict(space, t, z, tsp, tt, PRODUCT_SUM, [6, 9, 0.8, 0.7, 0.15, 0.10, 0.4], 1, 30, 3.0, 4,
                       f32_storage=f32)
        # ties are everywhere on a lattice: compare the candidate counts (filter parity) and
        # the untied rows only
        ctx_stats = ctx.kriging_stats()
        assert ctx_stats["candidates_in_window"] == int(o["n_cand"].sum())
        if (o["tie"] == 0).mean() > 0.9:
            compare(g, o, z, f32)


def test_model_eval(ctx):
    rng = np.random.default_rng(0)
    h = rng.uniform(0, 700, 1000)
    t = rng.uniform(0, 100, 1000)
    for models, params in [(SUM_METRIC, P_SUM_METRIC), (PRODUCT_SUM, P_PRODUCT_SUM),
                           (("sum_metric", "gaussian", "gaussian", "gaussian"), P_SUM_METRIC)]:
        ctx.kriging_set_model(*models, params)
        got = ctx.variogram_model_eval(h, t)
        want = ok.gamma(models, params, h, t)
        assert np.allclose(got, want, rtol=1e-13, atol=0)


def test_errors(ctx):
    c2 = _native.Context(0)
    z = np.zeros(4)
    with pytest.raises(ValueError):
        c2.kriging_predict_host(z, z, z, 1, 10, 1.0, 1.0)      # no model / obs
    c2.kriging_set_model(*SUM_METRIC, P_SUM_METRIC)
    with pytest.raises(ValueError):
        c2.kriging_predict_host(z, z, z, 1, 10, 1.0, 1.0)      # no observations
    c2.kriging_set_obs(z, z, z, z)
    with pytest.raises(ValueError):
        c2.kriging_predict_host(z, z, z, 1, 1000, 1.0, 1.0)    # k too large
    with pytest.raises(ValueError):
        c2.kriging_set_model("sum_metric", "spherical", "spherical", "spherical", [1, 2, 3])
    c2.close()


@pytest.mark.parametrize("solver", ["auto", "lu", "ldl"])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 32, 33, 46, 47, 48, 55, 62, 63, 64, 65, 96, 127,
                               128])
def test_every_kernel_variant(ctx, k, solver):
    """max_kriging_points selects the (warps per target, panel count) variant of the solve
    kernel and the padding of the system; sweep the boundaries."""
    space, t, z, tsp, tt = synth(30000, 96, seed=100 + k, extent=200.0)
    g = gpu_predict(ctx, space, t, z, tsp, tt, SUM_METRIC, P_SUM_METRIC, 1, k, 40.0, 40.0, f32=False,
                    solver=solver)
    o = ok.predict(space, t, z, tsp, tt, SUM_METRIC, P_SUM_METRIC, 1, k, 40.0, 40.0, f32_storage=False)
    assert o["n_cand"].max() > k
    compare(g, o, z, False)
    assert ctx.kriging_solver_stats()["cholesky_enabled"] == (solver == "auto")


def test_fewer_observations_than_k(ctx):
    space, t, z, tsp, tt = synth(12, 40, seed=8, extent=20.0)
    tt[:] = 400.0
    g = gpu_predict(ctx, space, t, z, tsp, tt, SUM_METRIC, P_SUM_METRIC, 1, 50, 1e3, 1e3, f32=False)
    o = ok.predict(space, t, z, tsp, tt, SUM_METRIC, P_SUM_METRIC, 1, 50, 1e3, 1e3, f32_storage=False)
    assert np.all(o["n_used"] == 12)
    compare(g, o, z, False)


@pytest.mark.parametrize("models,params,k", [
    (SUM_METRIC, P_SUM_METRIC, 20),
    (SUM_METRIC, [400, 60, 500, 0.4, 0.2, 0.3, 0.0, 0.0, 0.0, 5.0], 33),      # zero total nugget
    (PRODUCT_SUM, P_PRODUCT_SUM, 20),
    (PRODUCT_SUM, P_PRODUCT_SUM, 70),
])
def test_duplicate_observations_minimum_norm_parity(ctx, models, params, k):
    """Identical observations make [Gamma 1; 1' 0] exactly singular whatever the nugget (the
    reference puts gamma(0, 0) on the diagonal).  np.linalg.lstsq = dgelsd then returns the
    MINIMUM-NORM solution -- the copies of an observation share its weight equally -- provided
    its SVD puts the null singular values below rcond = eps; when rounding noise leaves one of
    them above eps (seen with many duplicates in one system), dgelsd divides by that noise and
    its answer carries an arbitrary null-space component.  The GPU path flags rank-deficient
    systems in the pivoted LU and re-solves them with a Jacobi SVD (solve_minnorm_kernel), which
    always returns the minimum-norm solution.  Checked here: (1) GPU == the pseudo-inverse
    solution built independently in numpy, for every rank-deficient target; (2) GPU == the
    oracle's dgelsd wherever dgelsd itself returned the minimum-norm solution.  The copies carry
    DIFFERENT residuals, so a wrong weight split shows in the mean."""
    space, t, z, tsp, tt = synth(2000, 60, seed=4, extent=60.0)
    space[1000:] = space[:1000]
    t[1000:] = t[:1000]                       # z[1000:] stays different from z[:1000]
    space[1900:] = space[:100]                # some points exist three times
    t[1900:] = t[:100]
    g = gpu_predict(ctx, space, t, z, tsp, tt, models, params, 1, k, 30.0, 30.0, f32=False)
    st = ctx.kriging_stats()
    assert st["targets_singular"] > 0
    o = ok.predict(space, t, z, tsp, tt, models, params, 1, k, 30.0, 30.0, f32_storage=False)
    rows = (o["tie"] == 0) & (o["n_used"] > 0)
    assert rows.sum() > 10
    assert np.array_equal(g["idx"][rows], o["idx"][rows])
    n_def = n_oracle_minnorm = 0
    for i in np.nonzero(rows)[0]:
        m = int(o["n_used"][i])
        ids = o["idx"][i][:m].astype(np.int64)
        xs, ys, ts = space[ids, 0], space[ids, 1], t[ids]
        h = np.sqrt((xs[:, None] - xs[None, :]) ** 2 + (ys[:, None] - ys[None, :]) ** 2)
        A = np.ones((m + 1, m + 1))
        A[:m, :m] = ok.gamma(models, params, h, np.abs(ts[:, None] - ts[None, :]))
        A[m, m] = 0.0
        dxs, dys = xs - tsp[i, 0], ys - tsp[i, 1]
        b = np.append(ok.gamma(models, params, np.sqrt(dxs * dxs + dys * dys), np.abs(ts - tt[i])), 1.0)
        sv = np.linalg.svd(A, compute_uv=False)
        if sv[-1] > 1e-12 * sv[0]:
            continue                           # full rank: covered by every other test
        n_def += 1
        w_ref = (np.linalg.pinv(A, rcond=1e-12, hermitian=True) @ b)[:m]
        wscale = np.abs(w_ref).max()
        assert np.all(np.abs(g["w"][i][:m] - w_ref) <= 2.0 ** -20 * wscale), i
        assert abs(g["mean"][i] - w_ref @ z[ids]) <= 1e-8 * np.sum(np.abs(w_ref * z[ids])), i
        # copies of one observation share their weight equally (the minimum-norm property)
        key = np.round(xs * 1e6) + 1e9 * np.round(ts * 1e3)
        for kk in np.unique(key):
            grp = np.nonzero(key == kk)[0]
            if len(grp) > 1:
                wg = g["w"][i][grp].astype(np.float64)
                assert np.all(np.abs(wg - wg.mean()) <= 1e-6 * wscale)
        if np.all(np.abs(o["w"][i][:m] - w_ref) <= 2.0 ** -20 * wscale):   # dgelsd found it too
            n_oracle_minnorm += 1
            assert abs(g["mean"][i] - o["mean"][i]) <= 1e-8 * np.sum(np.abs(w_ref * z[ids])), i
    assert n_def > 3 and st["targets_singular"] >= n_def
    print(f"rank-deficient targets {n_def}, of which dgelsd returned the minimum-norm solution for {n_oracle_minnorm}")


def test_non_finite_system_raises_linalg_error(ctx):
    """The reference's lstsq raises np.linalg.LinAlgError when the kriging matrix holds NaN / Inf
    (numba _check_finite_matrix, stif/predictor.py:587): STIF_ENONFINITE -> LinAlgError."""
    space, t, z, tsp, tt = synth(3000, 40, seed=6, extent=60.0)
    bad = [400, 60, 500, 0.4, 0.2, 0.3, 0.03, np.inf, 0.03, 5.0]      # infinite nugget: gamma = inf
    with pytest.raises(np.linalg.LinAlgError):
        gpu_predict(ctx, space, t, z, tsp, tt, SUM_METRIC, bad, 1, 12, 30.0, 30.0, f32=False)
    nanp = [300, 60, np.nan, 0.7, 0.15, 0.10, 0.4]
    with pytest.raises(np.linalg.LinAlgError):
        gpu_predict(ctx, space, t, z, tsp, tt, PRODUCT_SUM, nanp, 1, 12, 30.0, 30.0, f32=True)
    # and the context stays usable
    g = gpu_predict(ctx, space, t, z, tsp, tt, SUM_METRIC, P_SUM_METRIC, 1, 12, 30.0, 30.0, f32=False)
    assert np.all(np.isfinite(g["mean"]))


def test_device_numpy_row_sum_is_numpys(ctx):
    """The DEVICE summation code of the float32-storage epilogue (csrc/kriging.cu numpy_row_sum)
    against np.sum(axis=1), bit for bit, for every row length the kernel supports."""
    rng = np.random.default_rng(0)
    for k in range(1, 129):
        a64 = rng.normal(size=(96, k)) * 10.0 ** rng.uniform(-3, 3, size=(96, 1))
        a32 = a64.astype(np.float32)
        assert np.array_equal(ctx.debug_numpy_row_sum(a64), np.sum(a64, axis=1)), k
        assert np.array_equal(ctx.debug_numpy_row_sum(a32), np.sum(a32, axis=1)), k


def test_huge_radii_streaming_select(ctx):
    """Every observation is a candidate (tens of buffer refills per target)."""
    space, t, z, tsp, tt = synth(6000, 40, seed=12)
    tt[:] = 366.0
    g = gpu_predict(ctx, space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, 24, 1e4, 1e4, f32=False)
    o = ok.predict(space, t, z, tsp, tt, PRODUCT_SUM, P_PRODUCT_SUM, 1, 24, 1e4, 1e4, f32_storage=False)
    assert np.all(o["n_cand"] == 6000)
    compare(g, o, z, False)


def test_saturated_model_ties_broken_by_index(ctx):
    """Search radius far beyond the model ranges: most candidates sit at the sill, i.e. exact
    gamma ties at the top-k boundary -- the rule is lowest observation index."""
    space, t, z, tsp, tt = synth(4000, 60, seed=15, extent=100.0)
    params = [0.5, 0.5, 0.5, 0.4, 0.2, 0.3, 0.03, 0.03, 0.03, 1.0]   # ranges 0.5, radii 50
    g = gpu_predict(ctx, space, t, z, tsp, tt, SUM_METRIC, params, 1, 10, 50.0, 50.0, f32=False)
    o = ok.predict(space, t, z, tsp, tt, SUM_METRIC, params, 1, 10, 50.0, 50.0, f32_storage=False)
    assert o["tie"].mean() > 0.5
    # oracle applies the same (gamma, index) rule: the sets must agree even on tied rows
    assert np.array_equal(g["idx"], o["idx"])


@pytest.mark.parametrize("models,params,k", [(SUM_METRIC, P_SUM_METRIC, 50), (PRODUCT_SUM, P_PRODUCT_SUM, 50),
                                             (PRODUCT_SUM, P_PRODUCT_SUM, 100)])
def test_bitwise_reproducible(ctx, models, params, k):
    """Every output of the kriging path is a deterministic function of the inputs (no atomics on
    the data path; the fall-back list order does not matter): two calls agree bit for bit."""
    space, t, z, tsp, tt = synth(40000, 3000, seed=41)
    a = gpu_predict(ctx, space, t, z, tsp, tt, models, params, 1, k, 30.0, 30.0, f32=False)
    b = gpu_predict(ctx, space, t, z, tsp, tt, models, params, 1, k, 30.0, 30.0, f32=False)
    for key in ("mean", "std", "w", "gamma", "idx", "n_used"):
        assert np.array_equal(a[key], b[key], equal_nan=True), key


@pytest.mark.parametrize("models,params", [(SUM_METRIC, P_SUM_METRIC), (PRODUCT_SUM, P_PRODUCT_SUM),
                                           (("sum_metric", "gaussian", "spherical", "gaussian"), P_SUM_METRIC)])
def test_stand_alone_kriging_weights(ctx, models, params):
    """calc_kriging_weights (stif/predictor.py:566-593) as a call of its own -- the callable the
    reference installs as _kriging_weights_function -- against the oracle's dgelsd solve."""
    rng = np.random.default_rng(17)
    ctx.kriging_set_model(*models, params)
    for k in (1, 2, 7, 33, 100, 128):
        xs, ys, ts = rng.uniform(0, 40, k), rng.uniform(0, 40, k), rng.uniform(0, 40, k)
        g0 = ok.gamma(models, params, np.hypot(xs - 20.0, ys - 20.0), np.abs(ts - 41.0))
        w = ctx.kriging_weights(xs, ys, ts, g0)
        w_ref, rank = ok.weights(xs, ys, ts, g0, models, params)
        assert rank == k + 1
        assert abs(w.sum() - 1.0) < 1e-9
        assert np.all(np.abs(w - w_ref) <= 1e-9 * np.abs(w_ref).max()), k
    from stif_b200 import Data, Predictor
    import pandas as pd
    df = pd.DataFrame({"x": xs, "y": ys, "time": ts, "z": rng.normal(size=len(xs))})
    p = Predictor(Data(df, space_cols=["x", "y"], time_col="time", predictand_col="z"), context=ctx)
    p.set_variogram_model(params, *models)
    w2 = p._kriging_weights_function(g0, np.column_stack([xs, ys]), ts)     # the reference's signature
    assert np.array_equal(w2, w)
