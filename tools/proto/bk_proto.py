"""CPU prototype of the restricted-pivot Bunch-Kaufman LDL' on the covariance form (numerics + pivot statistics)."""
import sys
import numpy as np

ALPHA = (1.0 + 17.0 ** 0.5) / 8.0


def sph(h, r, c0, b):
    u = np.minimum(h / r, 1.0)
    return b + c0 * (1.5 * u - 0.5 * u ** 3)


def gamma_ps(h, t, p):
    rs, rt, cs, ct, bs, bt, k = p
    pr = sph(h, rs, cs, bs) * sph(t, rt, ct, bt)
    return k * pr + pr


def make_system(rng, p, kk, smax, tmax, ncand):
    r = smax * np.sqrt(rng.uniform(size=ncand))
    a = rng.uniform(0, 2 * np.pi, ncand)
    x, y = r * np.cos(a), r * np.sin(a)
    t = -rng.uniform(0, tmax, ncand)
    g0 = gamma_ps(np.hypot(x, y), np.abs(t), p)
    sel = np.argsort(g0, kind="stable")[:kk]
    x, y, t, g0 = x[sel], y[sel], t[sel], g0[sel]
    H = np.hypot(x[:, None] - x[None, :], y[:, None] - y[None, :])
    T = np.abs(t[:, None] - t[None, :])
    G = gamma_ps(H, T, p)
    return G, g0


def truth(G, g0):
    n = len(g0)
    A = np.ones((n + 1, n + 1), np.longdouble)
    A[:n, :n] = G
    A[n, n] = 0
    b = np.append(g0, 1.0).astype(np.longdouble)
    # GEPP in long double
    A = A.copy()
    N = n + 1
    for j in range(N):
        pv = j + np.argmax(np.abs(A[j:, j]))
        if pv != j:
            A[[j, pv]] = A[[pv, j]]
            b[[j, pv]] = b[[pv, j]]
        l = A[j + 1:, j] / A[j, j]
        A[j + 1:, j:] -= l[:, None] * A[j, j:][None, :]
        b[j + 1:] -= l * b[j]
    xs = np.zeros(N, np.longdouble)
    for j in range(N - 1, -1, -1):
        xs[j] = (b[j] - A[j, j + 1:] @ xs[j + 1:]) / A[j, j]
    return xs[:n].astype(np.float64)


def bk_solve(G, g0, csill, stats, pivoting=True):
    m = len(g0)
    n = m + 2
    A = np.zeros((n, n))
    A[:m, :m] = csill - G
    A[m, :m] = csill - g0
    A[:m, m] = A[m, :m]
    A[m + 1, :m] = 1.0
    A[:m, m + 1] = 1.0
    perm = np.arange(m)
    amax0 = np.abs(A[:m, :m]).max()
    growth = 1.0

    def swap(a, b):
        if a == b:
            return
        A[[a, b], :] = A[[b, a], :]
        A[:, [a, b]] = A[:, [b, a]]
        perm[[a, b]] = perm[[b, a]]

    j = 0
    while j < m:
        ajj = abs(A[j, j])
        col = np.abs(A[j + 1:m, j])
        if len(col):
            r = j + 1 + int(np.argmax(col))
            lam = col.max()
        else:
            r, lam = j, 0.0
        two = False
        if not pivoting or ajj >= ALPHA * lam:
            stats["p1"] += 1
        else:
            rowr = np.abs(A[r, j:m]).copy()
            rowr[r - j] = 0.0
            sig = rowr.max()
            if ajj * sig >= ALPHA * lam * lam:
                stats["p1s"] += 1
            elif abs(A[r, r]) >= ALPHA * sig:
                swap(j, r)
                stats["swap1"] += 1
            else:
                swap(j + 1, r)
                two = True
                stats["p2"] += 1
        if not two:
            d = A[j, j]
            l = A[j + 1:, j] / d
            A[j + 1:, j + 1:] -= np.outer(l, A[j + 1:, j])
            A[j + 1:, j] = l
            A[j, j + 1:] = l
            j += 1
        else:
            D = A[j:j + 2, j:j + 2].copy()
            W = A[j + 2:, j:j + 2].copy()
            det = D[0, 0] * D[1, 1] - D[0, 1] * D[0, 1]
            Linv = np.array([[D[1, 1], -D[0, 1]], [-D[0, 1], D[0, 0]]]) / det
            Lm = W @ Linv
            A[j + 2:, j + 2:] -= Lm @ W.T
            A[j + 2:, j:j + 2] = Lm
            A[j:j + 2, j + 2:] = Lm.T
            A[j + 1, j] = A[j, j + 1] = 0.0
            j += 2
        if j < m:
            growth = max(growth, np.abs(A[j:m, j:m]).max() / amax0)
    # z rows: A[m, :m] = z_v, A[m+1,:m] = z_u; corner Schur
    s_vu = A[m + 1, m]
    s_uu = A[m + 1, m + 1]
    lam_ = (1.0 + s_vu) / (-s_uu)
    t = A[m, :m] + lam_ * A[m + 1, :m]
    # back substitution L' w = t (unit lower L in A[:m,:m] strictly lower)
    w = t.copy()
    for i in range(m - 1, 0, -1):
        w[:i] -= A[i, :i] * w[i]
    out = np.zeros(m)
    out[perm] = w
    stats["growth"] = max(stats.get("growth", 1.0), growth)
    return out, growth


def main():
    kk = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    nsys = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    p = [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4]
    csill = (p[6] + 1) * (p[2] + p[4]) * (p[3] + p[5])
    rng = np.random.default_rng(5)
    stats = dict(p1=0, p1s=0, swap1=0, p2=0)
    stats_np = dict(p1=0, p1s=0, swap1=0, p2=0)
    errs, errs_np, errs_lstsq, gr, gr_np = [], [], [], [], []
    for s in range(nsys):
        G, g0 = make_system(rng, p, kk, 30.0, 30.0, 465 * kk // 100 if kk < 100 else 465)
        wt = truth(G, g0)
        sc = np.abs(wt).max()
        w, g = bk_solve(G, g0, csill, stats, True)
        w2, g2 = bk_solve(G, g0, csill, stats_np, False)
        n = len(g0)
        A = np.ones((n + 1, n + 1)); A[:n, :n] = G; A[n, n] = 0
        wl = np.linalg.lstsq(A, np.append(g0, 1.0), rcond=None)[0][:n]
        errs.append(np.abs(w - wt).max() / sc)
        errs_np.append(np.abs(w2 - wt).max() / sc)
        errs_lstsq.append(np.abs(wl - wt).max() / sc)
        gr.append(g); gr_np.append(g2)
    print("k", kk, "systems", nsys)
    print("BK     : max err %.2e  median %.2e  growth max %.1f" % (max(errs), np.median(errs), max(gr)), stats)
    print("no piv : max err %.2e  median %.2e  growth max %.1f" % (max(errs_np), np.median(errs_np), max(gr_np)))
    print("lstsq  : max err %.2e  median %.2e" % (max(errs_lstsq), np.median(errs_lstsq)))


if __name__ == "__main__":
    main()
