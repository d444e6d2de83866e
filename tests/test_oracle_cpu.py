"""CPU: the oracle (oracle/*.c) against the golden vectors of the unmodified reference.

This is what pins the oracle; the GPU tests then compare the CUDA path with the oracle
(and with the same golden vectors directly)."""
import os

import numpy as np
import pytest

from oracle import kriging as ok
from oracle import variogram as ov


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def bit_equal(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_variogram_pm10_bit_exact(golden_dir):
    d = load(golden_dir, "pm10.npz")
    g = load(golden_dir, "variogram_pm10.npz")
    space = np.column_stack([d["x"], d["y"]])
    vg, norm, ws, wt = ov.get_variogram(space, d["time"], d["pm10"], 6e5, 7, 10, 7)
    assert norm.sum() == 9009481          # SURVEY 8(c) known answer
    assert bit_equal(norm, g["norm"])
    assert bit_equal(vg, g["variogram"])   # same summation order -> identical doubles
    assert (ws, wt) == tuple(g["widths"])
    cg, cnorm, _, _ = ov.get_covariogram(space, d["time"], d["pm10"], 5e5, 10, 8, 5)
    assert bit_equal(cnorm, g["cnorm"])
    assert np.allclose(cg, g["covariogram"], rtol=1e-12, atol=0, equal_nan=True)
    vb, nb, _, _ = ov.get_variogram(space, d["time"], d["pm10"] > 20, 6e5, 7, 10, 7)
    assert bit_equal(nb, g["norm_binary"]) and bit_equal(vb, g["variogram_binary"])


def test_variogram_synth(golden_dir):
    g = load(golden_dir, "variogram_synth.npz")
    vg, norm, _, _ = ov.get_variogram(g["space"], g["t_int"], g["v"], 30.0, 30, 30, 30)
    assert bit_equal(norm, g["norm_int"]) and bit_equal(vg, g["vg_int"])
    vg, norm, _, _ = ov.get_variogram(g["space"], g["t_flt"], g["v"], 30.0, 30.0, 30, 30)
    assert bit_equal(norm, g["norm_flt"]) and bit_equal(vg, g["vg_flt"])
    vg, norm, _, _ = ov.get_variogram(np.asfortranarray(g["space"]), g["t_int"], g["vb"], 25.0, 12, 7, 5)
    assert bit_equal(norm, g["norm_bool"]) and bit_equal(vg, g["vg_bool"])
    cg, norm, _, _ = ov.get_covariogram(g["space"], g["t_flt"], g["v"] + 2.0, 30.0, 30.0, 9, 11)
    assert bit_equal(norm, g["cnorm_flt"])
    assert np.allclose(cg, g["cg_flt"], rtol=1e-11, atol=1e-15, equal_nan=True)


def test_variogram_codegen_fingerprint(golden_dir):
    """Adversarial 3-point inputs: the executed codegen (fma(dy,dy,dx*dx), reciprocal
    multiply, peeled first pair) reproduces the reference; each counter-factual does not."""
    g = load(golden_dir, "variogram_adversarial.npz")
    kinds = set()
    for i in range(len(g["kind"])):
        ns, nt = int(g["ns"][i]), int(g["nt"][i])
        space = np.column_stack([g["x"][i], g["y"][i]])
        want = g["norm"][i, : ns * nt].reshape(ns, nt)
        got = ov.raw(space, g["t"][i], g["v"], g["smax"][i], g["tmax"][i], ns, nt)[1]
        assert bit_equal(got, want), f"case {i} ({g['kind'][i]})"
        alt = ov.raw(space, g["t"][i], g["v"], g["smax"][i], g["tmax"][i], ns, nt,
                     flags=int(g["alt_flags"][i]))[1]
        assert not bit_equal(alt, want), f"case {i} does not discriminate {g['kind'][i]}"
        kinds.add(str(g["kind"][i]))
    assert kinds == {"d2_fma_x", "d2_unfused", "true_div", "no_peel"}


@pytest.mark.parametrize("name", ["c4", "c5", "gauss", "metric"])
def test_kriging_synth_bit_exact(golden_dir, name):
    g = load(golden_dir, "kriging_synth.npz")
    k, smax, tmax, min_pts = g[f"{name}_cfg"]
    o = ok.predict(g["space"], g["t"], g["z"], g["tsp"], g["tt"], tuple(g[f"{name}_models"]),
                   g[f"{name}_params"], int(min_pts), int(k), smax, tmax)
    assert o["tie"].sum() == 0
    assert bit_equal(o["idx"], g[f"{name}_idx"])           # neighbour sets and order
    assert bit_equal(o["gamma"], g[f"{name}_gamma"])
    # same LAPACK routine (dgelsd) -> float32-rounded weights identical
    assert np.mean(o["w"] == g[f"{name}_w"]) > 0.999
    assert np.allclose(o["w"], g[f"{name}_w"], rtol=1e-6, atol=1e-9)
    assert np.allclose(o["mean"], g[f"{name}_mean"], rtol=1e-7, atol=1e-9)
    assert np.allclose(o["std"], g[f"{name}_std"], rtol=1e-6, atol=1e-9, equal_nan=True)
    assert np.mean(o["mean"] == g[f"{name}_mean"]) > 0.99


@pytest.mark.parametrize("st", ["sum_metric", "product_sum"])
def test_kriging_pm10(golden_dir, st):
    d = load(golden_dir, "pm10.npz")
    g = load(golden_dir, "kriging_pm10.npz")
    space = np.column_stack([d["x"], d["y"]])
    sel = g["targets"]
    o = ok.predict(space, d["time"], d["pm10"], space[sel], d["time"][sel],
                   (st, "spherical", "spherical", "spherical"), g[f"{st}_params"], 1, int(g["k"]),
                   float(g["smax"]), float(g["tmax"]), leave_out_idxs=sel)
    untied = o["tie"] == 0
    assert untied.mean() > 0.5
    # same neighbour SETS on untied rows (the reference's unstable argsort may order
    # equal-gamma neighbours differently, which is allowed)
    for i in np.nonzero(untied)[0]:
        n = max(o["n_used"][i], 0)
        assert set(o["idx"][i][:n]) == set(g[f"{st}_idx"][i][:n])
    same_order = np.all(o["idx"] == g[f"{st}_idx"], axis=1)
    rows = untied & same_order
    assert rows.mean() > 0.5
    assert np.allclose(o["gamma"][rows], g[f"{st}_gamma"][rows], rtol=1e-6)
    if st == "product_sum":   # well-conditioned fit: same solver -> same numbers
        assert np.allclose(o["mean"][rows], g[f"{st}_mean"][rows], rtol=1e-6, atol=1e-6)


def test_numpy_row_sum_emulation():
    """The pairwise order the oracle and the CUDA epilogue emulate IS numpy's."""
    rng = np.random.default_rng(0)
    for k in (3, 7, 8, 10, 50, 100, 128):
        a64 = rng.normal(size=(64, k)) * 10 ** rng.uniform(-3, 3, size=(64, 1))
        a32 = a64.astype(np.float32)
        assert bit_equal([_pairwise(r, np.float64) for r in a64], np.sum(a64, axis=1))
        assert bit_equal([_pairwise(r, np.float32) for r in a32], np.sum(a32, axis=1))


def _pairwise(a, dt):
    n = len(a)
    if n < 8:
        res = dt(0)
        for v in a:
            res = dt(res + v)
    else:
        r = [dt(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = dt(r[j] + a[i + j])
            i += 8
        res = dt(dt(dt(r[0] + r[1]) + dt(r[2] + r[3])) + dt(dt(r[4] + r[5]) + dt(r[6] + r[7])))
        while i < n:
            res = dt(res + a[i])
            i += 1
    return dt(dt(0) + res)
