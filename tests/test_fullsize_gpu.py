"""GPU tests at BASELINE.json's full sizes (C2, C3, C4, C5), through the C ABI.

The oracle cannot run 5e9 .. 5e11 pairs or millions of kriging systems in seconds, so these
tests use size-independent properties of the domain

  * variogram: culled vs un-culled tile work lists (two independent paths through the pair
    kernel), invariance under a permutation of the observations, exact scaling of the sums
    when the values are multiplied by a power of two, partitions (GPUs) summing to the whole,
    and a closed-form total of the pair counts;
  * kriging: the weights sum to one, the two independent solvers (covariance-form Cholesky
    and pivoted LU of the bordered system) agree, targets are independent of how the call is
    batched, translation of all coordinates changes nothing,

plus a direct comparison with the CPU oracle on a bounded sample of the same workload (a row
block of the variogram triangle; a few hundred kriging targets against ALL observations).
Tolerances are the north star's: counts bit-exact, sums 1e-12, kriging 1e-9 relative.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, record_quantiles
from oracle import kriging as ok
from oracle import variogram as ov
from stif_b200 import _native

pytestmark = pytest.mark.gpu

VARIO = dict(smax=30.0, tmax=30.0, ns=30, nt=30)
C4_MODEL = ("sum_metric", "spherical", "spherical", "spherical")
C4_PARAMS = [400, 60, 500, 0.4, 0.2, 0.3, 0.03, 0.03, 0.03, 5.0]
C5_MODEL = ("product_sum", "spherical", "spherical", "spherical")
C5_PARAMS = [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4]


def synth_variogram(n):
    rng = np.random.default_rng(20261019)          # SURVEY 8(d) C2 / C3 generator
    x = rng.uniform(0, 100, n)
    y = rng.uniform(0, 100, n)
    t = rng.integers(0, 365, n).astype(np.float64)
    v = rng.normal(size=n)
    return x, y, t, v


def vario(ctx, x, y, t, v, flags=_native.VARIO_PEEL, **kw):
    return ctx.variogram_host(x, y, t, v, VARIO["smax"], VARIO["tmax"], VARIO["ns"], VARIO["nt"],
                              flags=flags, **kw)


def assert_sums_close(a, b, rtol=1e-12):
    assert np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), 1e-300))


@pytest.fixture(scope="module")
def c2(ctx):
    x, y, t, v = synth_variogram(100_000)
    counts, sums = vario(ctx, x, y, t, v)
    return (x, y, t, v), counts, sums


def test_c2_whole_workload_against_oracle_golden(ctx, c2):
    """All 4 999 950 000 pairs of C2 against the oracle's full run (oracle/gen_fullsize.py,
    tests/golden/variogram_c2.npz): counts bit-exact, sums 1e-12 relative."""
    (x, y, t, v), counts, sums = c2
    g = np.load(os.path.join(GOLDEN, "variogram_c2.npz"))
    assert int(g["n"]) == len(x) and int(g["pairs"]) == len(x) * (len(x) - 1) // 2
    assert np.array_equal(counts.astype(np.int64), g["counts"])
    assert int(counts.sum()) == 166_728_678
    assert_sums_close(sums, g["sums"])
    rel = np.abs(sums - g["sums"]) / np.abs(g["sums"])
    record_quantiles("variogram_c2_sums_vs_oracle", rel=rel.ravel())
    # Default mode: the FP64 sums are accumulated with L2 atomics, whose order depends on the
    # scheduling: repeatable to rounding only (bound asserted), the counts are integers.
    c_b, s_b = vario(ctx, x, y, t, v)
    assert np.array_equal(c_b, counts)
    assert_sums_close(s_b, sums, rtol=1e-13)
    # STIF_VARIO_DETERMINISTIC: fixed-point integer sums -- bit-identical from run to run, for any
    # work list (culled / un-culled / space-time sorted visit the pairs in different orders on
    # different blocks) and still within 1e-12 of the oracle
    det = _native.VARIO_PEEL | _native.VARIO_DETERMINISTIC
    runs = [vario(ctx, x, y, t, v, flags=det) for _ in range(3)]
    runs.append(vario(ctx, x, y, t, v, flags=det | _native.VARIO_SPACE_SORT))
    runs.append(vario(ctx, x, y, t, v, flags=det | _native.VARIO_NO_CULL))
    for c_d, s_d in runs:
        assert np.array_equal(c_d.astype(np.int64), g["counts"])
        assert np.array_equal(s_d, runs[0][1]), "deterministic mode: sums differ between runs / work lists"
    assert_sums_close(runs[0][1], g["sums"])
    # the observation order does not matter either (every term is the same integer)
    p = np.random.default_rng(7).permutation(len(x))
    c_p, s_p = vario(ctx, x[p], y[p], t[p], v[p], flags=_native.VARIO_DETERMINISTIC)
    c_q, s_q = vario(ctx, x, y, t, v, flags=_native.VARIO_DETERMINISTIC)
    assert np.array_equal(c_p, c_q) and np.array_equal(s_p, s_q)


def test_c2_culled_equals_unculled(ctx, c2):
    (x, y, t, v), counts, sums = c2
    st = ctx.variogram_stats()
    counts2, sums2 = vario(ctx, x, y, t, v, flags=_native.VARIO_PEEL | _native.VARIO_NO_CULL)
    st2 = ctx.variogram_stats()
    assert st2["pairs_evaluated"] > 3 * st["pairs_evaluated"]      # really a different work list
    assert np.array_equal(counts, counts2)
    assert_sums_close(sums, sums2)
    assert 0.030 < counts.sum() / (100_000 * 99_999 // 2) < 0.036  # 3.34 % binned
    # space-time sort mode (not the default at this size): a third work list, same histogram
    counts3, sums3 = vario(ctx, x, y, t, v, flags=_native.VARIO_PEEL | _native.VARIO_SPACE_SORT)
    st3 = ctx.variogram_stats()
    assert st3["pairs_evaluated"] != st["pairs_evaluated"]           # (not smaller at 100k)
    assert np.array_equal(counts, counts3)
    assert_sums_close(sums, sums3)


def test_c2_permutation_and_power_of_two_scaling(ctx, c2):
    (x, y, t, v), counts, sums = c2
    p = np.random.default_rng(1).permutation(len(x))
    # the peeled first pair (rows 0, 1 of the caller's arrays) is a property of the ORDER;
    # compare without it
    c_a, s_a = vario(ctx, x, y, t, v, flags=0)
    c_b, s_b = vario(ctx, x[p], y[p], t[p], v[p], flags=0)
    assert np.array_equal(c_a, c_b)
    assert_sums_close(s_a, s_b)
    c_c, s_c = vario(ctx, x, y, t, 4.0 * v, flags=0)
    assert np.array_equal(c_a, c_c)
    # every term scales exactly; only the (atomic) summation order may differ
    assert_sums_close(s_c, 16.0 * s_a)


def test_c2_row_block_against_oracle(ctx):
    """Rows 0..9999 of the C2 observations: 5e7 pairs the oracle does in a fraction of a second."""
    x, y, t, v = synth_variogram(100_000)
    n = 10_000
    counts, sums = vario(ctx, x[:n], y[:n], t[:n], v[:n])
    h, c = ov.raw(np.column_stack([x[:n], y[:n]]), t[:n], v[:n], VARIO["smax"], VARIO["tmax"], VARIO["ns"],
                  VARIO["nt"])
    assert np.array_equal(counts.astype(np.float64), c)
    assert_sums_close(sums, h)


def test_c2_space_cut_only_total_count(ctx):
    """With the time window wide open and one bin, the count is the number of pairs closer than
    smax (lag == smax lands in bin 1 of 1 and is dropped, utils.py:171-172) -- checked against
    an independent k-d tree count."""
    x, y, t, v = synth_variogram(100_000)
    smax = 2.0
    counts, _ = ctx.variogram_host(x, y, t, v, smax, 1e9, 1, 1, flags=0)
    # independent count: k-d tree (pairs counted in both orders, plus the n self pairs)
    from scipy.spatial import cKDTree
    tree = cKDTree(np.column_stack([x, y]))
    both = int(tree.count_neighbors(tree, smax * (1.0 - 1e-15)))
    assert int(counts.sum()) == (both - len(x)) // 2


def test_c3_partitions_and_permutation(ctx):
    """1M observations (5e11 pairs): 8 partitions (what 8 GPUs would each do) sum to the whole;
    a permuted input gives the same counts."""
    x, y, t, v = synth_variogram(1_000_000)
    # the whole 499 999 500 000-pair workload against the oracle's full run
    # (oracle/gen_fullsize.py, 8 host cores x 8.6 min; tests/golden/variogram_c3.npz)
    g = np.load(os.path.join(GOLDEN, "variogram_c3.npz"))
    c_g, s_g = vario(ctx, x, y, t, v)
    assert np.array_equal(c_g.astype(np.int64), g["counts"])
    assert int(c_g.sum()) == 16_704_985_684
    assert_sums_close(s_g, g["sums"])
    record_quantiles("variogram_c3_sums_vs_oracle", rel=(np.abs(s_g - g["sums"]) / np.abs(g["sums"])).ravel())
    counts, sums = vario(ctx, x, y, t, v, flags=0)
    pairs = 1_000_000 * 999_999 // 2
    assert 0.030 < counts.sum() / pairs < 0.036
    st = ctx.variogram_stats()
    # 1M observations run in the space-time sort mode by default; the time-sorted work list
    # visits about twice the pairs and must give the same histogram
    c_t, s_t = vario(ctx, x, y, t, v, flags=_native.VARIO_TIME_SORT)
    st_t = ctx.variogram_stats()
    assert st["pairs_evaluated"] < 0.7 * st_t["pairs_evaluated"]
    assert np.array_equal(c_t, counts)
    assert_sums_close(s_t, sums)
    acc_c = np.zeros_like(counts)
    acc_s = np.zeros_like(sums)
    for part in range(8):
        c, s = vario(ctx, x, y, t, v, part=part, nparts=8)        # PEEL: applied by partition 0
        acc_c += c
        acc_s += s
    assert np.array_equal(acc_c.astype(np.int64), g["counts"])   # what 8 GPUs sum to == the oracle
    assert_sums_close(acc_s, g["sums"])
    p = np.random.default_rng(2).permutation(len(x))
    c_p, s_p = vario(ctx, x[p], y[p], t[p], v[p], flags=0)
    assert np.array_equal(c_p, counts)
    assert_sums_close(s_p, sums)


# ---------------------------------------------------------------------------------------------
def synth_kriging(n_obs, n_t, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1000, n_obs)
    y = rng.uniform(0, 1000, n_obs)
    t = rng.uniform(0, 365, n_obs)
    z = np.sin(x / 150) + np.cos(y / 200) + 0.5 * np.sin(t / 30) + 0.3 * rng.normal(size=n_obs)
    tx = rng.uniform(0, 1000, n_t)
    ty = rng.uniform(0, 1000, n_t)
    tt = rng.uniform(30, 365, n_t)
    return (x, y, t, z), (tx, ty, tt)


def close_rel(a, b, scale, rtol=1e-9):
    return np.all(np.abs(a - b) <= rtol * scale)


def test_c4_solvers_agree_weights_sum_to_one_batching(ctx):
    (x, y, t, z), (tx, ty, tt) = synth_kriging(1_000_000, 200_000, 20261020)
    ctx.kriging_set_obs(x, y, t, z)
    ctx.kriging_set_model(*C4_MODEL, C4_PARAMS)
    m1, s1, w1, g1, i1, n1 = ctx.kriging_predict_host(tx, ty, tt, 1, 50, 30.0, 30.0, flags=0, want_weights=True)
    ss = ctx.kriging_solver_stats()
    st = ctx.kriging_stats()
    assert ss["cholesky_enabled"] and ss["cholesky_to_lu"] == 0 and st["targets_singular"] == 0
    assert 200 < st["candidates_in_window"] / len(tx) < 260          # C-bar ~ 226
    assert np.all(n1 == 50)
    assert np.all(np.abs(w1.astype(np.float64).sum(axis=1) - 1.0) < 1e-5)   # float32 outputs
    # the pivoted LU of the bordered system on the same targets
    m2, s2, w2, g2, i2, n2 = ctx.kriging_predict_host(tx, ty, tt, 1, 50, 30.0, 30.0,
                                                      flags=_native.KRIG_FORCE_LU, want_weights=True)
    assert np.array_equal(i1, i2) and np.array_equal(g1, g2)
    mscale = np.sum(np.abs(w1.astype(np.float64)) * np.abs(z[i1.astype(np.int64)]), axis=1)
    vscale = np.sum(np.abs(w1.astype(np.float64) * g1), axis=1)
    assert close_rel(m1, m2, mscale)
    assert close_rel(s1 ** 2, s2 ** 2, vscale)
    # a target's result does not depend on what else is in the call
    sel = slice(70_000, 70_512)
    m3, s3 = ctx.kriging_predict_host(tx[sel], ty[sel], tt[sel], 1, 50, 30.0, 30.0, flags=0)
    assert np.array_equal(m3, m1[sel]) and np.array_equal(s3, s1[sel])


def test_c4_sample_against_oracle(ctx):
    (x, y, t, z), (tx, ty, tt) = synth_kriging(1_000_000, 1000, 20261020)
    ctx.kriging_set_obs(x, y, t, z)
    ctx.kriging_set_model(*C4_MODEL, C4_PARAMS)
    space, tsp = np.column_stack([x, y]), np.column_stack([tx, ty])
    for f32 in (False, True):
        flags = _native.KRIG_F32_STORAGE if f32 else 0
        m, s, w, g, idx, nu = ctx.kriging_predict_host(tx, ty, tt, 1, 50, 30.0, 30.0, flags=flags,
                                                       want_weights=True)
        o = ok.predict(space, t, z, tsp, tt, C4_MODEL, C4_PARAMS, 1, 50, 30.0, 30.0, f32_storage=f32)
        rows = o["tie"] == 0
        assert rows.mean() > 0.99
        assert np.array_equal(idx[rows], o["idx"][rows])           # neighbour sets identical
        assert np.array_equal(nu[rows], o["n_used"][rows])
        mscale = np.sum(np.abs(o["w"].astype(np.float64)) * np.abs(z[o["idx"].astype(np.int64)]), axis=1)
        vscale = np.sum(np.abs(o["w"].astype(np.float64) * o["gamma"]), axis=1)
        tol = 2.0 ** -22 if f32 else 1e-9
        assert np.all(np.abs(m - o["mean"])[rows] <= tol * mscale[rows])
        assert np.all(np.abs(s ** 2 - o["std"] ** 2)[rows] <= 4 * tol * vscale[rows])
        # the north star's wording: PLAIN relative error of mean and variance
        rel_m = (np.abs(m - o["mean"]) / np.abs(o["mean"]))[rows]
        rel_v = (np.abs(s ** 2 - o["std"] ** 2) / np.abs(o["std"] ** 2))[rows]
        q = record_quantiles(f"c4_{'f32' if f32 else 'f64'}_vs_oracle", mean_rel=rel_m, var_rel=rel_v,
                             mean_over_scale=(np.abs(m - o["mean"]) / mscale)[rows])
        if not f32:
            assert q["mean_rel"]["p99"] <= 1e-9 and q["var_rel"]["max"] <= 1e-9
            # |mean| can be arbitrarily close to zero (weights of both signs times residuals of
            # both signs); the worst target is bounded through its cancellation factor instead
            assert q["mean_rel"]["max"] <= 1e-9 * float((mscale / np.abs(o["mean"]))[rows].max())
        else:
            assert q["mean_rel"]["p50"] <= 1e-9          # most float32 weights round the same way
            assert q["var_rel"]["max"] <= 2e-6           # float32 sum of float32 products


def test_c4_translation_invariance(ctx):
    (x, y, t, z), (tx, ty, tt) = synth_kriging(1_000_000, 20_000, 20261020)
    ctx.kriging_set_obs(x, y, t, z)
    ctx.kriging_set_model(*C4_MODEL, C4_PARAMS)
    m1, s1, w1, g1, i1, n1 = ctx.kriging_predict_host(tx, ty, tt, 1, 50, 30.0, 30.0, flags=0, want_weights=True)
    # power-of-two shifts keep every coordinate difference exact
    ctx.kriging_set_obs(x + 4096.0, y - 2048.0, t + 512.0, z)
    m2, s2, w2, g2, i2, n2 = ctx.kriging_predict_host(tx + 4096.0, ty - 2048.0, tt + 512.0, 1, 50, 30.0, 30.0,
                                                      flags=0, want_weights=True)
    same = np.all(i1 == i2, axis=1)
    assert same.mean() > 0.999          # differences are rounded at a different magnitude
    mscale = np.sum(np.abs(w1.astype(np.float64)) * np.abs(z[i1.astype(np.int64)]), axis=1)
    assert np.all(np.abs(m1 - m2)[same] <= 1e-7 * mscale[same])


def test_c5_indicator_kriging_k100(ctx):
    """2M observations, presence/absence residuals, product_sum, k = 100: the model is not a
    valid variogram, every system goes through the Bunch-Kaufman LDL' kernel; sample vs the oracle,
    all targets vs the pivoted LU."""
    rng = np.random.default_rng(20261021)
    n_obs = 2_000_000
    x = rng.uniform(0, 1000, n_obs)
    y = rng.uniform(0, 1000, n_obs)
    t = rng.uniform(0, 365, n_obs)
    field = np.sin(x / 150) + np.cos(y / 200) + 0.5 * np.sin(t / 30)
    presence = rng.uniform(size=n_obs) < 1.0 / (1.0 + np.exp(-2.0 * field))
    z = presence.astype(np.float64) - presence.mean()
    n_t = 20_000
    tx, ty, tt = rng.uniform(0, 1000, n_t), rng.uniform(0, 1000, n_t), rng.uniform(30, 365, n_t)
    ctx.kriging_set_obs(x, y, t, z)
    ctx.kriging_set_model(*C5_MODEL, C5_PARAMS)
    m, s, w, g, idx, nu = ctx.kriging_predict_host(tx, ty, tt, 1, 100, 30.0, 30.0, flags=0, want_weights=True)
    ss, st = ctx.kriging_solver_stats(), ctx.kriging_stats()
    # not a valid variogram: the symmetric indefinite LDL' kernel solves every system itself
    assert not ss["cholesky_enabled"] and ss["ldl_enabled"] and ss["first_pass_to_lu"] == 0
    assert 420 < st["candidates_in_window"] / n_t < 510               # C-bar ~ 465
    # the pivoted LU on all 20 000 targets: the same matrix, general elimination
    m_lu, s_lu, w_lu, _, idx_lu, _ = ctx.kriging_predict_host(tx, ty, tt, 1, 100, 30.0, 30.0,
                                                              flags=_native.KRIG_FORCE_LU, want_weights=True)
    assert not ctx.kriging_solver_stats()["ldl_enabled"]
    assert np.array_equal(idx, idx_lu)
    wsc = np.abs(w_lu).max(axis=1, keepdims=True)
    assert np.all(np.abs(w - w_lu) <= 2.0 ** -22 * wsc)
    msc = np.sum(np.abs(w_lu.astype(np.float64)) * np.abs(z[idx_lu.astype(np.int64)]), axis=1)
    record_quantiles("c5_f64_ldl_vs_lu", mean_over_scale=np.abs(m - m_lu) / msc)
    assert np.all(np.abs(m - m_lu) <= 1e-9 * msc)
    assert np.all(nu == 100)
    w64 = w.astype(np.float64)      # float32 outputs; product_sum weights are large and of both signs
    assert np.all(np.abs(w64.sum(axis=1) - 1.0) <= 1e-6 * np.abs(w64).sum(axis=1))
    assert np.isfinite(m).all()
    ns = 1000                      # oracle: dense filter over 2M observations + dgelsd of 101 x 101
    o = ok.predict(np.column_stack([x, y]), t, z, np.column_stack([tx[:ns], ty[:ns]]), tt[:ns], C5_MODEL,
                   C5_PARAMS, 1, 100, 30.0, 30.0, f32_storage=False)
    rows = o["tie"] == 0
    assert rows.mean() > 0.9
    assert np.array_equal(idx[:ns][rows], o["idx"][rows])
    mscale = np.sum(np.abs(o["w"].astype(np.float64)) * np.abs(z[o["idx"].astype(np.int64)]), axis=1)
    vscale = np.sum(np.abs(o["w"].astype(np.float64) * o["gamma"]), axis=1)
    assert np.all(np.abs(m[:ns] - o["mean"])[rows] <= 1e-9 * mscale[rows])
    # product_sum is not a valid variogram: sum(w gamma) is negative for part of the targets and
    # the reference's np.sqrt returns NaN there (predictor.py:905) -- same targets on both sides
    assert np.array_equal(np.isnan(s[:ns])[rows], np.isnan(o["std"])[rows])
    fin = rows & ~np.isnan(o["std"])
    assert fin.sum() > 0.5 * ns
    assert np.all(np.abs(s[:ns] ** 2 - o["std"] ** 2)[fin] <= 4e-9 * vscale[fin])
    rel_m = (np.abs(m[:ns] - o["mean"]) / np.abs(o["mean"]))[rows]
    rel_v = (np.abs(s[:ns] ** 2 - o["std"] ** 2) / np.abs(o["std"] ** 2))[fin]
    q = record_quantiles("c5_f64_vs_oracle", mean_rel=rel_m, var_rel=rel_v,
                         mean_over_scale=(np.abs(m[:ns] - o["mean"]) / mscale)[rows])
    assert q["mean_rel"]["p99"] <= 1e-9 * float(np.quantile((mscale / np.abs(o["mean"]))[rows], 0.99))
    # the stock float32-storage arithmetic on the same sample
    m32, s32 = ctx.kriging_predict_host(tx[:ns], ty[:ns], tt[:ns], 1, 100, 30.0, 30.0)
    o32 = ok.predict(np.column_stack([x, y]), t, z, np.column_stack([tx[:ns], ty[:ns]]), tt[:ns], C5_MODEL,
                     C5_PARAMS, 1, 100, 30.0, 30.0, f32_storage=True)
    rel_m32 = (np.abs(m32 - o32["mean"]) / np.abs(o32["mean"]))[rows]
    both = rows & ~np.isnan(s32) & ~np.isnan(o32["std"])
    rel_s32 = (np.abs(s32 - o32["std"]) / np.abs(o32["std"]))[both]
    record_quantiles("c5_f32_vs_oracle", mean_rel=rel_m32, std_rel=rel_s32,
                     mean_over_scale=(np.abs(m32 - o32["mean"]) / mscale)[rows])
    assert np.all(np.abs(m32 - o32["mean"])[rows] <= 2.0 ** -21 * mscale[rows])


def test_pipelined_host_call_equals_plain_calls(ctx):
    """Host calls with more than 600k targets take the chunked two-stream pipeline (copies of
    the neighbouring chunks under the kernels): bit-identical to plain calls on pieces below the
    threshold, statistics accumulated over the chunks, optional outputs and leave-out included."""
    (x, y, t, z), (tx, ty, tt) = synth_kriging(1_000_000, 1_500_000, 77)
    ctx.kriging_set_obs(x, y, t, z)
    ctx.kriging_set_model(*C4_MODEL, C4_PARAMS)
    m, s = ctx.kriging_predict_host(tx, ty, tt, 1, 50, 30.0, 30.0)
    st = ctx.kriging_stats()
    ph = ctx.last_phase_ms()
    assert ph[1] > 0 and ph[2] > 0
    pieces = [slice(0, 500_000), slice(500_000, 1_000_000), slice(1_000_000, 1_500_000)]
    ms, ss, sts = [], [], []
    for p in pieces:
        a, b = ctx.kriging_predict_host(tx[p], ty[p], tt[p], 1, 50, 30.0, 30.0)
        ms.append(a)
        ss.append(b)
        sts.append(ctx.kriging_stats())
    assert np.array_equal(m, np.concatenate(ms)) and np.array_equal(s, np.concatenate(ss))
    for key in st:
        assert st[key] == sum(q[key] for q in sts), key
    # optional outputs + leave-out
    n = 900_000
    lo = np.random.default_rng(3).integers(0, len(x), n)
    a = ctx.kriging_predict_host(tx[:n], ty[:n], tt[:n], 1, 8, 30.0, 30.0, leave_out=lo, want_weights=True)
    h = 450_000
    b1 = ctx.kriging_predict_host(tx[:h], ty[:h], tt[:h], 1, 8, 30.0, 30.0, leave_out=lo[:h], want_weights=True)
    b2 = ctx.kriging_predict_host(tx[h:n], ty[h:n], tt[h:n], 1, 8, 30.0, 30.0, leave_out=lo[h:], want_weights=True)
    for q in range(6):
        assert np.array_equal(a[q], np.concatenate([b1[q], b2[q]])), q
