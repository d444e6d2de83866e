"""Generate tests/golden/*.npz from the UNMODIFIED reference (numba path).

Run in the build container only (needs /root/reference):   python -m oracle.gen_golden
The fixtures pin the CPU oracle (oracle/*.c) and, through it, the CUDA path.  Nothing
here is used at test time; the tests read the committed .npz files.

Fixtures
  pm10.npz                   the reference's PM10 Germany test table (tests/data/pm10.csv)
  variogram_pm10.npz         all-pairs variogram + covariogram of PM10 (reference output)
  variogram_synth.npz        synthetic inputs (int64 and float64 time, bool values) + output
  variogram_adversarial.npz  3-point inputs that discriminate the executed fast-math
                             codegen (FMA form of d^2, reciprocal multiply, peeled pair)
  kriging_pm10.npz           kriging (w, gamma, idx, mean, std) on PM10 targets, fitted
                             sum_metric and product_sum models, k=50, with leave-out
  kriging_synth.npz          synthetic kriging case, several models
"""
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402
from oracle import variogram as ov  # noqa: E402
from oracle import kriging as ok  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore")


def ref_predictor(stif, space, t, z, models=None, params=None):
    df = pd.DataFrame({"x": space[:, 0], "y": space[:, 1], "time": t, "z": z})
    data = stif.Data(df, space_cols=["x", "y"], time_col="time", predictand_col="z",
                     covariate_cols=[])
    p = stif.Predictor(data, None)
    p._prepare_geostatistics()
    if models is not None:
        inject(p, models, params)
    return p


def inject(p, models, params):
    """Install a fixed parameter vector instead of running Nelder-Mead."""
    p._variogram_fit = types.SimpleNamespace(x=np.asarray(params, dtype=float))
    p._variogram_models = list(models)
    p._variogram_model_function = p._create_variogram_model_function()
    p._kriging_weights_function = p._create_kriging_weights_function()
    p._kriging_function = p._create_kriging_function()


def gen_pm10(stif):
    df = refload.load_pm10()
    np.savez_compressed(os.path.join(GOLDEN, "pm10.npz"), x=df["x"].to_numpy(), y=df["y"].to_numpy(),
                        time=df["time"].to_numpy(), pm10=df["PM10"].to_numpy())
    return df


def gen_variogram_pm10(stif, df):
    from stif.utils import get_covariogram, get_variogram
    space = df[["x", "y"]].to_numpy()
    t = df["time"].to_numpy()
    v = df["PM10"].to_numpy()
    vg, norm, ws, wt = get_variogram(space, t, v, 6e5, 7, 10, 7, None)
    cg, cnorm, cws, cwt = get_covariogram(space, t, v, 5e5, 10, 8, 5, None)
    vb, nb, _, _ = get_variogram(space, t, v > 20, 6e5, 7, 10, 7, None)   # df_binary of the tests
    np.savez_compressed(os.path.join(GOLDEN, "variogram_pm10.npz"),
                        variogram=vg, norm=norm, widths=np.array([ws, wt]),
                        params=np.array([6e5, 7, 10, 7]),
                        covariogram=cg, cnorm=cnorm, cparams=np.array([5e5, 10, 8, 5]),
                        variogram_binary=vb, norm_binary=nb)
    print("PM10 variogram: total pairs binned", norm.sum(), "vg[0,0]", vg[0, 0], "vg[9,6]", vg[9, 6])


def gen_variogram_synth(stif):
    from stif.utils import get_covariogram, get_variogram
    rng = np.random.default_rng(20261019)
    n = 3000
    space = rng.uniform(0, 100, (n, 2))
    t_int = rng.integers(0, 365, n)
    t_flt = rng.uniform(0, 365, n)
    v = rng.normal(size=n)
    vb = rng.uniform(size=n) > 0.4
    out = dict(space=space, t_int=t_int, t_flt=t_flt, v=v, vb=vb)
    vg, norm, _, _ = get_variogram(space, t_int, v, 30.0, 30, 30, 30, None)       # C2 parameters
    out.update(vg_int=vg, norm_int=norm)
    vg, norm, _, _ = get_variogram(space, t_flt, v, 30.0, 30.0, 30, 30, None)
    out.update(vg_flt=vg, norm_flt=norm)
    vg, norm, _, _ = get_variogram(np.asfortranarray(space), t_int, vb, 25.0, 12, 7, 5, None)
    out.update(vg_bool=vg, norm_bool=norm)
    cg, norm, _, _ = get_covariogram(space, t_flt, v + 2.0, 30.0, 30.0, 9, 11, None)
    out.update(cg_flt=cg, cnorm_flt=norm)
    np.savez_compressed(os.path.join(GOLDEN, "variogram_synth.npz"), **out)


def gen_variogram_adversarial(stif):
    """Search 3-point inputs where the counter-factual codegens disagree with the executed
    one, and record what the reference does on them."""
    from stif.utils import get_variogram
    rng = np.random.default_rng(1)
    v = np.array([0.0, 1.0, 3.0])
    want = {"d2_fma_x": ov.FLAG_PEEL | ov.FLAG_D2_FMA_X, "d2_unfused": ov.FLAG_PEEL | ov.FLAG_D2_UNFUSED,
            "true_div": ov.FLAG_PEEL | ov.FLAG_TRUE_DIV, "no_peel": 0}
    cases = []
    found = {k: 0 for k in want}
    trials = 0
    while min(found.values()) < 8 and trials < 2_000_000:
        trials += 1
        x = rng.uniform(0, 100, 3)
        y = rng.uniform(0, 100, 3)
        float_time = bool(trials & 2)
        t = rng.uniform(0, 50, 3) if float_time else np.floor(rng.uniform(0, 50, 3))
        ns, nt = int(rng.integers(2, 40)), int(rng.integers(2, 12))
        first_pair = bool(trials & 1)
        a, b = (0, 1) if first_pair else (1, 2)
        far = 2 if first_pair else 0
        t[far] = 1e6
        lag = np.hypot(x[a] - x[b], y[a] - y[b])
        lag_t = abs(t[a] - t[b])
        k = int(rng.integers(1, ns))
        smax = lag * ns / k * (1 + int(rng.integers(-2, 3)) * 2.2e-16)
        if trials % 3 == 0 and lag_t > 0:
            kt = int(rng.integers(1, nt))
            tmax = lag_t * nt / kt * (1 + int(rng.integers(-2, 3)) * 2.2e-16)
        else:
            tmax = float(rng.integers(3, 30))
        space = np.column_stack([x, y])
        base = ov.raw(space, t, v, smax, tmax, ns, nt, flags=ov.REFERENCE_FLAGS)[1]
        for name, fl in want.items():
            if found[name] >= 8:
                continue
            if (name == "no_peel") != first_pair:
                continue
            alt = ov.raw(space, t, v, smax, tmax, ns, nt, flags=fl)[1]
            if np.array_equal(alt, base):
                continue
            tt = t if float_time else t.astype(np.int64)
            _, norm, _, _ = get_variogram(space, tt, v, smax, tmax, ns, nt, None)
            cases.append(dict(kind=name, x=x, y=y, t=t.copy(), float_time=float_time, smax=smax,
                              tmax=tmax, ns=ns, nt=nt, norm=norm, alt_flags=fl))
            found[name] += 1
            break
    print("adversarial cases:", found, "after", trials, "trials")
    maxb = 40 * 12
    K = len(cases)
    norm_flat = np.zeros((K, maxb))
    for i, c in enumerate(cases):
        norm_flat[i, : c["ns"] * c["nt"]] = c["norm"].ravel()
    np.savez_compressed(
        os.path.join(GOLDEN, "variogram_adversarial.npz"),
        kind=np.array([c["kind"] for c in cases]), x=np.array([c["x"] for c in cases]),
        y=np.array([c["y"] for c in cases]), t=np.array([c["t"] for c in cases]),
        float_time=np.array([c["float_time"] for c in cases]),
        smax=np.array([c["smax"] for c in cases]), tmax=np.array([c["tmax"] for c in cases]),
        ns=np.array([c["ns"] for c in cases]), nt=np.array([c["nt"] for c in cases]),
        alt_flags=np.array([c["alt_flags"] for c in cases]), norm=norm_flat, v=v)


def run_kriging(p, tsp, tt, k, smax, tmax, leave, min_pts=1):
    w, kv, idx = p._get_kriging_weights(tsp, tt, min_pts, k, smax, tmax, leave)
    mean, std = p.get_kriging_prediction(tsp, tt, min_pts, k, smax, tmax, leave)
    return dict(w=w, gamma=kv, idx=idx, mean=mean, std=std)


def gen_kriging_pm10(stif, df):
    space = df[["x", "y"]].to_numpy()
    t = df["time"].to_numpy()
    z = df["PM10"].to_numpy()
    p = ref_predictor(stif, space, t, z)
    # fit once with the real reference to obtain realistic parameter vectors
    p.calc_empirical_variogram(space_dist_max=6e5, time_dist_max=7, n_space_bins=10, n_time_bins=7)
    out = {}
    rng = np.random.default_rng(5)
    sel = np.sort(rng.choice(len(t), 400, replace=False))
    out["targets"] = sel
    for st in ("sum_metric", "product_sum"):
        p.fit_variogram_model(st_model=st)
        params = np.array(p._variogram_fit.x)
        print(st, "fit fun", p._variogram_fit.fun, "x", params)
        inject(p, (st, "spherical", "spherical", "spherical"), params)
        r = run_kriging(p, space[sel], t[sel], 50, 2e5, 5, sel)
        tie = ok.predict(space, t, z, space[sel], t[sel], (st, "spherical", "spherical", "spherical"),
                         params, 1, 50, 2e5, 5, leave_out_idxs=sel)["tie"]
        out[f"{st}_params"] = params
        for key, val in r.items():
            out[f"{st}_{key}"] = val
        out[f"{st}_tie"] = tie
    out["k"] = 50
    out["smax"] = 2e5
    out["tmax"] = 5
    np.savez_compressed(os.path.join(GOLDEN, "kriging_pm10.npz"), **out)


def gen_kriging_synth(stif):
    rng = np.random.default_rng(20261020)
    n = 6000
    space = rng.uniform(0, 160, (n, 2))
    t = rng.uniform(0, 365, n)
    z = np.sin(space[:, 0] / 150) + np.cos(space[:, 1] / 200) + 0.5 * np.sin(t / 30) + 0.3 * rng.normal(size=n)
    T = 300
    tsp = rng.uniform(0, 160, (T, 2))
    tt = rng.uniform(30, 365, T)
    out = dict(space=space, t=t, z=z, tsp=tsp, tt=tt)
    cases = [
        ("c4", ("sum_metric", "spherical", "spherical", "spherical"),
         [400, 60, 500, 0.4, 0.2, 0.3, 0.03, 0.03, 0.03, 5.0], 50, 30.0, 30.0, 1),
        ("c5", ("product_sum", "spherical", "spherical", "spherical"),
         [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4], 100, 45.0, 45.0, 1),
        ("gauss", ("sum_metric", "gaussian", "spherical", "gaussian"),
         [400, 60, 500, 0.4, 0.2, 0.3, 0.03, 0.03, 0.03, 5.0], 10, 30.0, 30.0, 5),
        ("metric", ("metric", "spherical", "spherical", "spherical"), [500, 1.0, 0.05, 5.0], 16, 30.0, 30.0, 1),
    ]
    for name, models, params, k, smax, tmax, min_pts in cases:
        p = ref_predictor(stif, space, t, z, models, params)
        r = run_kriging(p, tsp, tt, k, smax, tmax, None, min_pts)
        o = ok.predict(space, t, z, tsp, tt, models, params, min_pts, k, smax, tmax)
        out[f"{name}_models"] = np.array(models)
        out[f"{name}_params"] = np.array(params, dtype=float)
        out[f"{name}_cfg"] = np.array([k, smax, tmax, min_pts], dtype=float)
        out[f"{name}_tie"] = o["tie"]
        for key, val in r.items():
            out[f"{name}_{key}"] = val
        print(name, "oracle==reference: idx", np.array_equal(o["idx"], r["idx"]), "w",
              np.array_equal(o["w"], r["w"]), "mean", np.array_equal(o["mean"], r["mean"]))
    np.savez_compressed(os.path.join(GOLDEN, "kriging_synth.npz"), **out)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    stif = refload.load()
    df = gen_pm10(stif)
    gen_variogram_pm10(stif, df)
    gen_variogram_synth(stif)
    gen_variogram_adversarial(stif)
    gen_kriging_synth(stif)
    gen_kriging_pm10(stif, df)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
