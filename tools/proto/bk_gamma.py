"""BK on the bordered gamma system (the reference's own matrix), RHS riding along as an extra row."""
import sys
import numpy as np
sys.path.insert(0, "tools/proto")
from bk_proto import make_system, truth, ALPHA


def bk_gamma(G, g0, stats, alpha=ALPHA):
    m = len(g0)
    N = m + 1              # pivot rows
    n = N + 1              # + rhs row
    A = np.zeros((n, n))
    A[:m, :m] = G
    A[m, :m] = 1.0
    A[:m, m] = 1.0
    A[N, :m] = g0
    A[:m, N] = g0
    A[N, m] = A[m, N] = 1.0
    perm = np.arange(N)

    def swap(a, b):
        if a == b:
            return
        A[[a, b], :] = A[[b, a], :]
        A[:, [a, b]] = A[:, [b, a]]
        perm[[a, b]] = perm[[b, a]]

    j = 0
    while j < N:
        ajj = abs(A[j, j])
        col = np.abs(A[j + 1:N, j])
        if len(col):
            r = j + 1 + int(np.argmax(col)); lam = col.max()
        else:
            r, lam = j, 0.0
        two = False
        if ajj >= alpha * lam:
            stats["p1"] += 1
        else:
            rowr = np.abs(A[r, j:N]).copy(); rowr[r - j] = 0.0
            sig = rowr.max()
            if ajj * sig >= alpha * lam * lam:
                stats["p1s"] += 1
            elif abs(A[r, r]) >= alpha * sig:
                swap(j, r); stats["swap1"] += 1
            else:
                if r != j + 1:
                    stats["swap2"] += 1
                swap(j + 1, r); two = True; stats["p2"] += 1
        if not two:
            d = A[j, j]
            l = A[j + 1:, j] / d
            A[j + 1:, j + 1:] -= np.outer(l, A[j + 1:, j])
            A[j + 1:, j] = l; A[j, j + 1:] = l
            j += 1
        else:
            D = A[j:j + 2, j:j + 2].copy(); W = A[j + 2:, j:j + 2].copy()
            det = D[0, 0] * D[1, 1] - D[0, 1] * D[0, 1]
            Linv = np.array([[D[1, 1], -D[0, 1]], [-D[0, 1], D[0, 0]]]) / det
            Lm = W @ Linv
            A[j + 2:, j + 2:] -= Lm @ W.T
            A[j + 2:, j:j + 2] = Lm; A[j:j + 2, j + 2:] = Lm.T
            A[j + 1, j] = A[j, j + 1] = 0.0
            j += 2
    w = A[N, :N].copy()
    for i in range(N - 1, 0, -1):
        w[:i] -= A[i, :i] * w[i]
    out = np.zeros(N)
    out[perm] = w
    return out[:m]


if __name__ == "__main__":
    kk = int(sys.argv[1]); nsys = int(sys.argv[2])
    p = [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4]
    rng = np.random.default_rng(5)
    st = dict(p1=0, p1s=0, swap1=0, p2=0, swap2=0)
    errs = []
    for s in range(nsys):
        G, g0 = make_system(rng, p, kk, 30.0, 30.0, 465 * kk // 100 if kk < 100 else 465)
        wt = truth(G, g0); sc = np.abs(wt).max()
        errs.append(np.abs(bk_gamma(G, g0, st) - wt).max() / sc)
    print("bk_gamma k", kk, "max %.2e median %.2e" % (max(errs), np.median(errs)), st)
