"""Lane-level emulation of the planned kernel: packed lower storage, row ownership lane + 32 g."""
import sys
import numpy as np
sys.path.insert(0, "tools/proto")
from bk_proto import make_system, truth, ALPHA


def T(i):
    return i * (i + 1) // 2


def trunc_hi(x, bits):
    """|x| with the low `bits` bits of the high word (and the whole low word) cleared."""
    u = np.abs(np.float64(x)).view(np.uint64)
    hi = (int(u) >> 32) & ~((1 << bits) - 1)
    return np.uint64(hi << 32).view(np.float64)


def solve(G, g0, stats, tol_rel=1e-10):
    m = len(g0); N = m + 1; n = m + 2
    NG = (n + 31) // 32
    A = np.full(T(n) + 8, np.nan)
    for a in range(m):
        for b in range(a + 1):
            A[T(a) + b] = G[a, b]
    for e in range(m):
        A[T(m) + e] = 1.0
        A[T(m + 1) + e] = g0[e]
    A[T(m) + m] = 0.0
    A[T(m + 1) + m] = 1.0
    A[T(m + 1) + m + 1] = 0.0
    amax = max(np.abs(G).max(), 1.0)
    tol = tol_rel * amax
    perm = list(range(N))
    lanes = range(32)

    def do_swap(p, r):
        pairs = []
        for g in range(NG):
            for lane in lanes:
                i = lane + 32 * g
                if i >= n:
                    continue
                if i < p: o1, o2 = T(p) + i, T(r) + i
                elif i == p: o1, o2 = T(p) + p, T(r) + r
                elif i < r: o1, o2 = T(i) + p, T(r) + i
                elif i == r: continue
                else: o1, o2 = T(i) + p, T(i) + r
                pairs.append((o1, o2))
        flat = [o for pr in pairs for o in pr]
        assert len(flat) == len(set(flat))
        for o1, o2 in pairs:
            A[o1], A[o2] = A[o2], A[o1]
        perm[p], perm[r] = perm[r], perm[p]

    j = 0
    while j < N:
        v = {}
        key = 0
        for g in range(NG):
            for lane in lanes:
                i = lane + 32 * g
                v[i] = A[T(i) + j] if (i > j and i < n) else 0.0
                if i > j and i < N:
                    kk = (int(np.abs(np.float64(v[i])).view(np.uint64)) >> 32) & ~0xff | i
                    key = max(key, kk)
        ajj = A[T(j) + j]
        r = key & 0xff
        lam = np.uint64((key & ~0xff) << 32).view(np.float64)
        aj = abs(ajj)
        kind = 0
        if aj >= ALPHA * lam:
            stats["p1"] += 1
        else:
            sk = 0.0
            for c in range(j, N):
                if c == r: continue
                val = A[T(r) + c] if c < r else A[T(c) + r]
                sk = max(sk, trunc_hi(val, 0))
            sig = sk
            arr = abs(A[T(r) + r])
            if aj * sig >= ALPHA * lam * lam: kind = 0; stats["p1s"] += 1
            elif arr >= ALPHA * sig: kind = 1; stats["swap1"] += 1
            else: kind = 2; stats["p2"] += 1
        if kind == 1:
            do_swap(j, r)
        if kind == 2 and r != j + 1:
            do_swap(j + 1, r)
        if kind < 2:
            d = A[T(j) + j]
            if not abs(d) > tol:
                return None
            dinv = 1.0 / d
            w1 = {i: (A[T(i) + j] if j < i < n else 0.0) for i in range(32 * NG)}
            for i in range(j + 1, n):
                A[T(i) + j] = w1[i] * dinv
            for c in range(j + 1, n):
                l1 = A[T(c) + j]
                for i in range(c, n):
                    A[T(i) + c] -= w1[i] * l1
            j += 1
        else:
            if not lam > tol:
                return None
            d11, d21, d22 = A[T(j) + j], A[T(j + 1) + j], A[T(j + 1) + j + 1]
            t = 1.0 / d21
            ak, akp1 = d11 * t, d22 * t
            den = t / (ak * akp1 - 1.0)
            w1 = {i: (A[T(i) + j] if j + 1 < i < n else 0.0) for i in range(32 * NG)}
            w2 = {i: (A[T(i) + j + 1] if j + 1 < i < n else 0.0) for i in range(32 * NG)}
            for i in range(j + 2, n):
                A[T(i) + j] = den * (akp1 * w1[i] - w2[i])
                A[T(i) + j + 1] = den * (ak * w2[i] - w1[i])
            A[T(j + 1) + j] = 0.0
            for c in range(j + 2, n):
                l1, l2 = A[T(c) + j], A[T(c) + j + 1]
                for i in range(c, n):
                    A[T(i) + c] -= w1[i] * l1 + w2[i] * l2
            j += 2
    z = [A[T(N) + c] for c in range(N)]
    for i in range(N - 1, 0, -1):
        xi = z[i]
        for c in range(i):
            z[c] -= A[T(i) + c] * xi
    sol = np.zeros(m)
    for c in range(N):
        if perm[c] < m:
            sol[perm[c]] = z[c]
    return sol


if __name__ == "__main__":
    kk = int(sys.argv[1]); nsys = int(sys.argv[2])
    p = [300, 60, 0.8, 0.7, 0.15, 0.10, 0.4]
    rng = np.random.default_rng(5)
    st = dict(p1=0, p1s=0, swap1=0, p2=0)
    errs = []
    for s in range(nsys):
        G, g0 = make_system(rng, p, kk, 30.0, 30.0, 465 * kk // 100 if kk < 100 else 465)
        wt = truth(G, g0); sc = np.abs(wt).max()
        w = solve(G, g0, st)
        errs.append(np.abs(w - wt).max() / sc)
    print("packed k", kk, "max %.2e median %.2e" % (max(errs), np.median(errs)), st)
