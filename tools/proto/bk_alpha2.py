import sys
import numpy as np
sys.path.insert(0, "tools/proto")
from bk_proto import make_system, truth
from bk_gamma import bk_gamma
kk, nsys = 60, 25
for name, p in [("c5", [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4]), ("short ranges", [40, 20, 0.8, 0.7, 0.15, 0.10, 0.4]),
                ("zero nugget", [300, 60, 0.8, 0.7, 0.0, 0.0, 0.4]), ("tiny nugget", [300, 60, 0.8, 0.7, 1e-6, 1e-6, 2.0]),
                ("negative k", [300, 60, 0.8, 0.7, 0.15, 0.10, -0.5]), ("big nugget", [300, 60, 0.1, 0.1, 2.0, 3.0, 0.4])]:
    for alpha in (0.6403882032022076, 0.3):
        rng = np.random.default_rng(11)
        st = dict(p1=0, p1s=0, swap1=0, p2=0, swap2=0)
        errs, conds = [], []
        for s in range(nsys):
            G, g0 = make_system(rng, p, kk, 30.0, 30.0, 280)
            wt = truth(G, g0); sc = np.abs(wt).max()
            errs.append(np.abs(bk_gamma(G, g0, st, alpha) - wt).max() / sc)
            n = len(g0); A = np.ones((n + 1, n + 1)); A[:n, :n] = G; A[n, n] = 0
            lu = np.linalg.solve(A, np.append(g0, 1.0))[:n]
            conds.append(np.abs(lu - wt).max() / sc)
        steps = st["p1"] + st["p1s"] + st["swap1"] + st["p2"]
        print("%-13s alpha %.2f  BK max err %.2e  (LU %.2e)  fast %.0f%%" % (name, alpha, max(errs), max(conds), 100 * st["p1"] / steps))
