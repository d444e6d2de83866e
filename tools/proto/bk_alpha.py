"""How the Bunch-Kaufman threshold alpha changes the share of cheap pivots (no second scan) and the error."""
import sys
import numpy as np
sys.path.insert(0, "tools/proto")
from bk_proto import make_system, truth
from bk_gamma import bk_gamma

kk = int(sys.argv[1]); nsys = int(sys.argv[2])
p = [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4]
for alpha in (0.6403882032022076, 0.5, 0.4, 0.3, 0.2):
    rng = np.random.default_rng(5)
    st = dict(p1=0, p1s=0, swap1=0, p2=0, swap2=0)
    errs = []
    for s in range(nsys):
        G, g0 = make_system(rng, p, kk, 30.0, 30.0, 465 * kk // 100 if kk < 100 else 465)
        wt = truth(G, g0); sc = np.abs(wt).max()
        errs.append(np.abs(bk_gamma(G, g0, st, alpha) - wt).max() / sc)
    steps = st["p1"] + st["p1s"] + st["swap1"] + st["p2"]
    print("alpha %.3f  max err %.2e median %.2e  fast %.1f%%  scan-no-swap %.1f%% swap %.1f%% 2x2 %.1f%% (steps %d)" % (
        alpha, max(errs), np.median(errs), 100 * st["p1"] / steps, 100 * st["p1s"] / steps, 100 * st["swap1"] / steps,
        100 * st["p2"] / steps, steps))
