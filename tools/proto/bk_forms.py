import sys
import numpy as np
import scipy.linalg as sl
sys.path.insert(0, "tools/proto")
from bk_proto import make_system, truth, bk_solve

kk = int(sys.argv[1]); nsys = int(sys.argv[2])
p = [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4]
csill = (p[6] + 1) * (p[2] + p[4]) * (p[3] + p[5])
rng = np.random.default_rng(5)
E = {k: [] for k in ["bk_cov", "bk_cov_c2", "lu_cov", "lu_gamma", "sytrf_gamma", "sytrf_cov", "cond"]}
for s in range(nsys):
    G, g0 = make_system(rng, p, kk, 30.0, 30.0, 465 * kk // 100 if kk < 100 else 465)
    wt = truth(G, g0); sc = np.abs(wt).max(); n = len(g0)
    st = dict(p1=0, p1s=0, swap1=0, p2=0)
    E["bk_cov"].append(np.abs(bk_solve(G, g0, csill, st)[0] - wt).max() / sc)
    E["bk_cov_c2"].append(np.abs(bk_solve(G, g0, 0.5 * csill, st)[0] - wt).max() / sc)
    # LU on covariance form: solve [C -1; 1' 0][w; lam] = [c0; 1]
    C = csill - G
    K = np.zeros((n + 1, n + 1)); K[:n, :n] = C; K[:n, n] = -1; K[n, :n] = 1
    w = np.linalg.solve(K, np.append(csill - g0, 1.0))[:n]
    E["lu_cov"].append(np.abs(w - wt).max() / sc)
    A = np.ones((n + 1, n + 1)); A[:n, :n] = G; A[n, n] = 0
    w = np.linalg.solve(A, np.append(g0, 1.0))[:n]
    E["lu_gamma"].append(np.abs(w - wt).max() / sc)
    w = sl.solve(A, np.append(g0, 1.0), assume_a="sym")[:n]
    E["sytrf_gamma"].append(np.abs(w - wt).max() / sc)
    Ks = np.zeros((n + 1, n + 1)); Ks[:n, :n] = C; Ks[:n, n] = 1; Ks[n, :n] = 1
    w = sl.solve(Ks, np.append(csill - g0, 1.0), assume_a="sym")[:n]
    E["sytrf_cov"].append(np.abs(w - wt).max() / sc)
    E["cond"].append(np.linalg.cond(A))
for k, v in E.items():
    print("%-12s max %.2e median %.2e" % (k, max(v), np.median(v)))
