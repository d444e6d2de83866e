"""Emulation of the blocked kernel: per panel of up to 8 columns the matrix is only touched by swaps; the
column under examination (and row r of the Bunch-Kaufman test) is brought up to date on the fly from
the panel's finished columns (L stored in place, W = L D recomputed from D), and the trailing matrix
receives  -L W'  once per panel (the tensor-core part of the kernel)."""
import sys
import numpy as np
sys.path.insert(0, "tools/proto")
from bk_proto import make_system, truth, ALPHA
from bk_packed import T, trunc_hi

NB = 8


def solve(G, g0, stats, tol_rel=1e-10):
    m = len(g0); N = m + 1; n = m + 2
    A = np.full(T(n) + 64, 1e300)          # finite garbage outside
    for a in range(m):
        for b in range(a + 1):
            A[T(a) + b] = G[a, b]
    for e in range(m):
        A[T(m) + e] = 1.0
        A[T(m + 1) + e] = g0[e]
    A[T(m) + m] = 0.0; A[T(m + 1) + m] = 1.0; A[T(m + 1) + m + 1] = 0.0
    amax = max(np.abs(G).max(), 1.0)
    tol = tol_rel * amax
    perm = list(range(N))

    def do_swap(p, r):
        for i in range(n):
            if i < p: o1, o2 = T(p) + i, T(r) + i
            elif i == p: o1, o2 = T(p) + p, T(r) + r
            elif i < r: o1, o2 = T(i) + p, T(r) + i
            elif i == r: continue
            else: o1, o2 = T(i) + p, T(i) + r
            A[o1], A[o2] = A[o2], A[o1]
        perm[p], perm[r] = perm[r], perm[p]

    j0 = 0
    while j0 < N:
        dd = np.zeros(NB + 1); dl = np.zeros(NB + 1); du = np.zeros(NB + 1)

        def wrow(x, q):
            """W(x, 0..q-1) = (L D)(x, .) from row x's stored multipliers."""
            w = np.zeros(NB)
            for qq in range(q):
                w[qq] = dd[qq] * A[T(x) + j0 + qq]
                if dl[qq] != 0.0: w[qq] += dl[qq] * A[T(x) + j0 + qq - 1]
                if du[qq] != 0.0: w[qq] += du[qq] * A[T(x) + j0 + qq + 1]
            return w

        def upd_col(x, q, lo):
            """updated entries (i, x), i in [lo, n), as a dict; x >= j."""
            w = wrow(x, q)
            out = {}
            for i in range(lo, n):
                val = A[T(i) + x] if i >= x else A[T(x) + i]
                for qq in range(q):
                    val -= A[T(i) + j0 + qq] * w[qq]
                out[i] = val
            return out

        j = j0
        forced = 0
        while j < N and j - j0 < NB:
            q = j - j0
            col = upd_col(j, q, j)                 # rows j .. n-1 (own rows + the diagonal)
            d = col[j]
            if forced == 0:
                key = 0
                for i in range(j + 1, N):
                    kk = (int(np.abs(np.float64(col[i])).view(np.uint64)) >> 32) & ~0xff | i
                    key = max(key, kk)
                r = key & 0xff
                lam = np.uint64((key & ~0xff) << 32).view(np.float64)
                aj = abs(d)
                kind = 0
                if aj >= ALPHA * lam:
                    stats["p1"] += 1
                else:
                    rowr = upd_col(r, q, j)         # symmetric: entries (r, c) / (c, r), c in [j, n)
                    sk = lam
                    for c in range(j + 1, N):
                        if c != r: sk = max(sk, trunc_hi(rowr[c], 0))
                    arr = abs(rowr[r])
                    sig = sk
                    if aj * sig >= ALPHA * lam * lam: kind = 0; stats["p1s"] += 1
                    elif arr >= ALPHA * sig: kind = 1; stats["swap1"] += 1
                    else: kind = 2; stats["p2"] += 1
                if kind == 2 and q + 1 >= NB:
                    stats["straddle"] += 1
                    break
                if kind == 1:
                    do_swap(j, r); forced = 1; continue
                if kind == 2 and r != j + 1:
                    do_swap(j + 1, r); forced = 2; lam_keep = lam; continue
                if kind == 2: lam_keep = lam
            else:
                kind = 1 if forced == 1 else 2
                forced = 0
                lam = lam_keep if kind == 2 else 0.0
            if kind < 2:
                if not abs(d) > tol: return None
                dinv = 1.0 / d
                for i in range(j + 1, n): A[T(i) + j] = col[i] * dinv
                dd[q] = d
                j += 1
            else:
                if not lam > tol: return None
                col2 = upd_col(j + 1, q, j + 1)
                d11, d21, d22 = d, col[j + 1], col2[j + 1]
                t = 1.0 / d21
                ak, akp1 = d11 * t, d22 * t
                den = t / (ak * akp1 - 1.0)
                for i in range(j + 2, n):
                    A[T(i) + j] = den * (akp1 * col[i] - col2[i])
                    A[T(i) + j + 1] = den * (ak * col2[i] - col[i])
                A[T(j + 1) + j] = 0.0
                dd[q] = d11; du[q] = d21; dd[q + 1] = d22; dl[q + 1] = d21
                j += 2
        j1 = j
        pw = j1 - j0
        # trailing update, tiles of the absolute 8-grid with masks (here entry by entry)
        for c in range(j1, n):
            w = wrow(c, pw)
            for i in range(c, n):
                s = 0.0
                for qq in range(pw): s += A[T(i) + j0 + qq] * w[qq]
                A[T(i) + c] -= s
        j0 = j1
    z = [A[T(N) + c] for c in range(N)]
    for i in range(N - 1, 0, -1):
        xi = z[i]
        for c in range(i): z[c] -= A[T(i) + c] * xi
    sol = np.zeros(m)
    for c in range(N):
        if perm[c] < m: sol[perm[c]] = z[c]
    return sol


if __name__ == "__main__":
    kk = int(sys.argv[1]); nsys = int(sys.argv[2])
    p = [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4]
    rng = np.random.default_rng(5)
    st = dict(p1=0, p1s=0, swap1=0, p2=0, straddle=0)
    errs = []
    for s in range(nsys):
        G, g0 = make_system(rng, p, kk, 30.0, 30.0, 465 * kk // 100 if kk < 100 else 465)
        wt = truth(G, g0); sc = np.abs(wt).max()
        w = solve(G, g0, st)
        errs.append(np.abs(w - wt).max() / sc)
    print("blocked k", kk, "max %.2e median %.2e" % (max(errs), np.median(errs)), st)
