"""GPU parity: libstif_b200 variogram kernels vs the CPU oracle (bit-exact counts,
sums within 1e-12 relative) -- through the C ABI (stif_variogram_host)."""
import numpy as np
import pytest

from oracle import variogram as ov
from stif_b200 import _native

pytestmark = pytest.mark.gpu

RTOL_SUMS = 1e-12  # north-star tolerance for the binned semivariance sums


def synth(n, seed=20261019, int_time=True):
    rng = np.random.default_rng(seed)
    space = rng.uniform(0, 100, (n, 2))
    t = rng.integers(0, 365, n) if int_time else rng.uniform(0, 365, n)
    v = rng.normal(size=n)
    return space, t, v


def check(ctx, space, t, v, smax, tmax, ns, nt, mode=0, flags=_native.VARIO_PEEL):
    x, y = _native.split_space(space)
    mean = float(np.mean(np.asarray(v, dtype=np.float64))) if mode == 1 else 0.0
    counts, sums = ctx.variogram_host(x, y, t, v, smax, tmax, ns, nt, mode=mode, mean=mean, flags=flags)
    h, c = ov.raw(space, t, v, smax, tmax, ns, nt, mode=mode)
    assert np.array_equal(counts.astype(np.float64), c), "bin pair counts must be bit-exact"
    if not flags & (_native.VARIO_NO_CULL | _native.VARIO_SPACE_SORT):
        # the space-time sort mode (default from 160k observations) is held to the same bar
        check(ctx, space, t, v, smax, tmax, ns, nt, mode=mode, flags=flags | _native.VARIO_SPACE_SORT)
    scale = np.maximum(np.abs(h), 1e-300)
    if mode == 0:
        assert np.all(np.abs(sums - h) <= RTOL_SUMS * scale)
    else:
        # covariogram sums cancel; bound by the sum of |terms| ~ counts * var
        bound = RTOL_SUMS * np.maximum(c * np.var(v), 1e-300)
        assert np.all(np.abs(sums - h) <= bound)
    return counts, sums


@pytest.mark.parametrize("n", [2, 3, 255, 256, 257, 1023, 1024, 1025, 2500, 5000])
def test_sizes(ctx, n):
    space, t, v = synth(n, seed=n)
    check(ctx, space, t, v, 30.0, 30, 30, 30)


def test_empty_and_single(ctx):
    for n in (0, 1):
        counts, sums = ctx.variogram_host(np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n), 1.0, 1.0, 3, 4)
        assert counts.shape == (3, 4) and counts.sum() == 0 and np.all(sums == 0)


@pytest.mark.parametrize("int_time", [True, False])
@pytest.mark.parametrize("flags", [_native.VARIO_PEEL, _native.VARIO_PEEL | _native.VARIO_NO_CULL])
def test_synth_20k(ctx, int_time, flags):
    space, t, v = synth(20000, int_time=int_time)
    check(ctx, space, t, v, 30.0, 30, 30, 30, flags=flags)


def test_covariogram(ctx):
    space, t, v = synth(12000, seed=5)
    check(ctx, space, t, v + 3.0, 25.0, 20.0, 12, 9, mode=1)


def test_all_pairs_in_one_bin(ctx):
    space, t, v = synth(4000, seed=9)
    counts, _ = check(ctx, space, t, v, 1e3, 1e3, 1, 1)
    assert counts.sum() == 4000 * 3999 // 2


def test_duplicates_and_edges(ctx):
    # integer lattice: many exact ties on bin edges and on the cut-offs (lag == max)
    g = np.arange(40, dtype=np.float64)
    xx, yy = np.meshgrid(g, g)
    space = np.column_stack([xx.ravel(), yy.ravel()])
    rng = np.random.default_rng(3)
    t = rng.integers(0, 12, len(space))
    v = rng.normal(size=len(space))
    check(ctx, space, t, v, 10.0, 5, 10, 5)
    check(ctx, space, t, v, 10.0, 5, 7, 3)


@pytest.mark.parametrize("offset", [0.0, 5.4e6, -3.3e9, 1e15])
def test_float_prefilter_large_offsets(ctx, offset):
    """The sweep's float pre-filter centres the coordinates and keeps every pair within its
    error bound of the cut-offs; pairs are decided in FP64.  Large common offsets (UTM-like
    metres, 1e15) leave few or no float bits for the differences: counts must stay bit-exact.
    Lattice spacing 1 with cut-offs on lattice distances puts many pairs exactly ON them."""
    g = np.arange(48, dtype=np.float64)
    xx, yy = np.meshgrid(g, g)
    space = np.column_stack([xx.ravel() + offset, yy.ravel() - 2 * offset])
    rng = np.random.default_rng(5)
    t = rng.integers(0, 20, len(space)).astype(np.float64) + offset
    v = rng.normal(size=len(space))
    check(ctx, space, t, v, 5.0, 5.0, 5, 5)      # 3-4-5 triangles: lag == max, dropped
    check(ctx, space, t, v, 13.0, 3.0, 13, 3)    # 5-12-13


def test_float_prefilter_overflow_and_nan(ctx):
    """Coordinates beyond the float range (the filter passes everything, the drain decides)
    and NaN coordinates / times (never bin, as in the oracle)."""
    rng = np.random.default_rng(6)
    n = 3000
    space = rng.uniform(0, 50, (n, 2))
    t = rng.uniform(0, 30, n)
    v = rng.normal(size=n)
    space[::7, 0] *= 1e40
    space[5::11, 1] = -1e300
    t[3::13] = 1e200
    check(ctx, space, t, v, 10.0, 5.0, 10, 5)
    space[1::17, 0] = np.nan
    t[2::19] = np.nan
    x, y = _native.split_space(space)
    counts, sums = ctx.variogram_host(x, y, t, v, 10.0, 5.0, 10, 5)
    ok = np.isfinite(space).all(axis=1) & np.isfinite(t)
    x2, y2 = _native.split_space(space[ok])
    c2, s2 = ctx.variogram_host(x2, y2, t[ok], v[ok], 10.0, 5.0, 10, 5, flags=0)
    c1, _ = ctx.variogram_host(x, y, t, v, 10.0, 5.0, 10, 5, flags=0)
    assert np.array_equal(c1, c2)


def test_float_prefilter_tiny_radius(ctx):
    """Cut-offs far below the float resolution of the centred coordinates: the error bound
    dominates the radius, the filter keeps a wide band, the result is still exact."""
    rng = np.random.default_rng(8)
    n = 5000
    base = rng.integers(0, 40, (n, 2)).astype(np.float64) * 1e6
    space = base + rng.uniform(0, 1e-3, (n, 2))
    t = rng.integers(0, 3, n)
    v = rng.normal(size=n)
    check(ctx, space, t, v, 1e-3, 2, 8, 2)


@pytest.mark.parametrize("t0,span,tmax", [(1.7e9, 40000.0, 700.0), (-3.0e5, 900.0, 0.05), (0.0, 1e-300, 1e-301)])
def test_coarse_time_sort(ctx, t0, span, tmax):
    """The observations are sorted by the top 32 bits of the time key only; tiles are bounded
    by coarse-bucket edges.  Times whose differences live far below that resolution (unix
    seconds with a minutes window, a sub-resolution window, a window at 1e-301) and more than
    one i-tile: culling and the skipped time test must not lose or add a pair."""
    rng = np.random.default_rng(17)
    n = 5000
    space = rng.uniform(0, 20, (n, 2))
    t = t0 + rng.uniform(0, span, n)
    t[::97] = t0            # exact ties, and for t0 = 0 a mix of +0.0 and -0.0
    t[5::193] = -t[5::193] if t0 == 0.0 else t[5::193]
    if t0 == 0.0:
        t[::194] = -0.0
    v = rng.normal(size=n)
    counts, _ = check(ctx, space, t, v, 8.0, tmax, 8, 6)
    assert counts.sum() > 0


def test_unsorted_negative_times(ctx):
    rng = np.random.default_rng(11)
    n = 6000
    space = rng.uniform(-50, 50, (n, 2))
    t = rng.uniform(-1e3, 1e3, n)
    v = rng.normal(size=n) * 100
    check(ctx, space, t, v, 40.0, 150.0, 16, 16)


def test_bool_values_f_order(ctx):
    rng = np.random.default_rng(12)
    n = 5000
    space = np.asfortranarray(rng.uniform(0, 10, (n, 2)))
    t = rng.integers(0, 50, n)
    v = rng.uniform(size=n) > 0.5
    check(ctx, space, t, v, 3.0, 10, 10, 10)


def test_partitions_sum_to_whole(ctx):
    space, t, v = synth(9000, seed=21)
    x, y = _native.split_space(space)
    whole_c, whole_s = ctx.variogram_host(x, y, t, v, 30.0, 30, 30, 30)
    acc_c = np.zeros_like(whole_c)
    acc_s = np.zeros_like(whole_s)
    for p in range(3):
        c, s = ctx.variogram_host(x, y, t, v, 30.0, 30, 30, 30, part=p, nparts=3)
        acc_c += c
        acc_s += s
    assert np.array_equal(acc_c, whole_c)
    assert np.allclose(acc_s, whole_s, rtol=1e-13, atol=0)


def test_stats_and_culling(ctx):
    space, t, v = synth(30000)
    x, y = _native.split_space(space)
    c1, _ = ctx.variogram_host(x, y, t, v, 30.0, 30, 30, 30)
    st_cull = ctx.variogram_stats()
    c2, _ = ctx.variogram_host(x, y, t, v, 30.0, 30, 30, 30, flags=_native.VARIO_PEEL | _native.VARIO_NO_CULL)
    st_all = ctx.variogram_stats()
    assert np.array_equal(c1, c2)
    assert st_cull["pairs"] == 30000 * 29999 // 2
    assert st_cull["pairs_in_window"] == st_all["pairs_in_window"]
    assert st_cull["pairs_evaluated"] < st_all["pairs_evaluated"]


@pytest.mark.parametrize("ns,nt", [(8192, 1), (1, 8192), (90, 91), (1000, 8)])
def test_many_bins(ctx, ns, nt):
    """Bin grids up to the 8192-bin limit: shared-memory sizing of the pair kernel, the float
    guard of the space-bin lookup at its widest (2^-19 * 8192), thresholds for 8192 bins."""
    space, t, v = synth(3000, seed=41, int_time=False)
    counts, _ = check(ctx, space, t, v, 35.0, 40.0, ns, nt)
    assert counts.sum() > 0
    # the deterministic mode keeps four histograms per block in shared memory (131 KB at 8192 bins)
    x, y = _native.split_space(space)
    det = _native.VARIO_PEEL | _native.VARIO_DETERMINISTIC
    if 36 * 1024 + 8 * (ns + 2) + 16 * ns * nt > 227 * 1024:       # does not fit one block: a clear error
        with pytest.raises(ValueError, match="STIF_VARIO_DETERMINISTIC"):
            ctx.variogram_host(x, y, t, v, 35.0, 40.0, ns, nt, flags=det)
        return
    c_d, s_d = ctx.variogram_host(x, y, t, v, 35.0, 40.0, ns, nt, flags=det)
    c_e, s_e = ctx.variogram_host(x, y, t, v, 35.0, 40.0, ns, nt, flags=det)
    assert np.array_equal(c_d, counts) and np.array_equal(s_d, s_e)


@pytest.mark.parametrize("smax", [1e-310, 5e-324, 1e308])
def test_degenerate_bin_widths(ctx, smax):
    """Space cut-offs whose bin width or its reciprocal leaves the float / double range: the
    kernel falls back to the reference's expression per binned pair (no threshold table)."""
    rng = np.random.default_rng(43)
    n = 1500
    space = rng.integers(0, 3, (n, 2)).astype(np.float64) * (smax if smax < 1 else 1.0)
    t = rng.integers(0, 5, n)
    v = rng.normal(size=n)
    check(ctx, space, t, v, smax, 3, 4, 3)


def test_bad_arguments(ctx):
    z = np.zeros(4)
    with pytest.raises(ValueError):
        ctx.variogram_host(z, z, z, z, 1.0, 1.0, 0, 3)
    with pytest.raises(ValueError):
        ctx.variogram_host(z, z, z, z, 1.0, 1.0, 200, 200)
    with pytest.raises(ValueError):
        ctx.variogram_host(z, z, z, z, 1.0, 1.0, 3, 3, part=2, nparts=2)


def _both_sort_modes(ctx, x, y, t, v, smax, tmax, ns, nt):
    a, sa = ctx.variogram_host(x, y, t, v, smax, tmax, ns, nt, flags=_native.VARIO_PEEL | _native.VARIO_TIME_SORT)
    b, sb = ctx.variogram_host(x, y, t, v, smax, tmax, ns, nt, flags=_native.VARIO_PEEL | _native.VARIO_SPACE_SORT)
    assert np.array_equal(a, b), "space-time sort mode changes the bin counts"
    assert np.allclose(sa, sb, rtol=1e-12, atol=0.0)
    return a


def test_space_sort_clustered_and_degenerate_extents(ctx):
    """Space-time sort mode: clustered stations (many tiles culled in space), a single time
    stamp, a single location, one coordinate constant, windows larger than the data."""
    rng = np.random.default_rng(77)
    n = 30000
    centres = rng.uniform(0, 1000, (12, 2))
    k = rng.integers(0, 12, n)
    x = centres[k, 0] + rng.normal(scale=4.0, size=n)
    y = centres[k, 1] + rng.normal(scale=4.0, size=n)
    t = rng.uniform(0, 2000, n)
    v = rng.normal(size=n)
    c = _both_sort_modes(ctx, x, y, t, v, 25.0, 40.0, 12, 9)
    assert c.sum() > 0
    space = np.stack([x, y], axis=1)[:6000]
    check(ctx, space, t[:6000], v[:6000], 25.0, 40.0, 12, 9)
    _both_sort_modes(ctx, x, y, np.full(n, 17.0), v, 25.0, 40.0, 12, 9)            # one time stamp
    _both_sort_modes(ctx, np.full(n, 3.0), np.full(n, -8.0), t, v, 25.0, 40.0, 12, 9)  # one location
    _both_sort_modes(ctx, x, np.zeros(n), t, v, 25.0, 40.0, 12, 9)                 # y constant
    _both_sort_modes(ctx, x, y, t, v, 1e9, 1e9, 7, 5)                               # nothing to cull
    _both_sort_modes(ctx, x, y, t, v, 25.0, 1e-3, 12, 9)                            # > 2^16 slabs


def test_space_sort_nan_inf(ctx):
    """NaN / infinite coordinates and times in the space-time sort mode: same counts as the
    time-sorted mode (which is checked against the oracle elsewhere)."""
    rng = np.random.default_rng(78)
    n = 9000
    x = rng.uniform(0, 100, n); y = rng.uniform(0, 100, n)
    t = rng.uniform(0, 100, n); v = rng.normal(size=n)
    for arr in (x, y, t):
        arr[rng.integers(0, n, 40)] = np.nan
        arr[rng.integers(0, n, 5)] = np.inf
        arr[rng.integers(0, n, 5)] = -np.inf
    _both_sort_modes(ctx, x, y, t, v, 20.0, 15.0, 10, 8)
    x[:] = np.nan
    c = _both_sort_modes(ctx, x, y, t, v, 20.0, 15.0, 10, 8)
    assert c.sum() == 0


@pytest.mark.parametrize("mode", [0, 1])
def test_deterministic_mode(ctx, mode):
    """STIF_VARIO_DETERMINISTIC: every term enters an integer histogram as a 62-bit fixed-point
    number (quantum q = 2^-60 .. 2^-59 of the largest possible term R), so the sums are
    bit-identical between runs and work lists; a bin's sum is within count * q / 2 of the exact
    sum of its terms -- i.e. 1e-12 relative for any bin whose mean term exceeds 1e-6 R, and an
    ABSOLUTE 2^-60 R otherwise (sparse bins of nearly equal values keep the default mode's
    relative accuracy only in the default mode)."""
    space, t, v = synth(6000, seed=77)
    v = v * 3.0 + (5.0 if mode == 1 else 0.0)
    x, y = _native.split_space(space)
    mean = float(np.mean(v)) if mode == 1 else 0.0
    h, c = ov.raw(space, t, v, 30.0, 30, 30, 30, mode=mode)
    outs = []
    for flags in (_native.VARIO_PEEL | _native.VARIO_DETERMINISTIC,
                  _native.VARIO_PEEL | _native.VARIO_DETERMINISTIC,
                  _native.VARIO_PEEL | _native.VARIO_DETERMINISTIC | _native.VARIO_NO_CULL,
                  _native.VARIO_PEEL | _native.VARIO_DETERMINISTIC | _native.VARIO_SPACE_SORT):
        outs.append(ctx.variogram_host(x, y, t, v, 30.0, 30, 30, 30, mode=mode, mean=mean, flags=flags))
    for counts, sums in outs:
        assert np.array_equal(counts.astype(np.float64), c)
        assert np.array_equal(sums, outs[0][1])                  # bit-identical
    R = (v.max() - v.min()) ** 2 if mode == 0 else np.max(np.abs(v - mean)) ** 2
    q = R * 2.0 ** -59
    bound = RTOL_SUMS * (np.abs(h) if mode == 0 else c * np.var(v)) + c * q
    assert np.all(np.abs(outs[0][1] - h) <= bound)
    # non-finite values poison their bins (NaN), like an FP64 sum would, and only those
    v2 = v.copy()
    v2[17] = np.nan
    _, s_nan = ctx.variogram_host(x, y, t, v2, 30.0, 30, 30, 30, mode=0, flags=_native.VARIO_DETERMINISTIC)
    h2, _ = ov.raw(space, t, v2, 30.0, 30, 30, 30, mode=0, flags=0)
    assert np.array_equal(np.isnan(s_nan), np.isnan(h2))
