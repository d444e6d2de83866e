// krig_ldl.cuh -- symmetric indefinite kriging solve: Bunch-Kaufman LDL' of the reference's own
// bordered system, one warp per target.  Included by kriging.cu (namespace stif::krig).
//
// Replaces calc_kriging_weights (stif/predictor.py:566-593) for models that are NOT valid
// variograms: product / product_sum (stif/variogram_models.py:126-127, :155-157) give an
// indefinite gamma matrix, unconstrained Nelder-Mead fits a negative total nugget.  The reference
// solves  K x = b,  K = [Gamma 1; 1' 0],  b = [gamma0; 1]  with LAPACK dgelsd; K is symmetric, so
//      P K P' = L D L'      (D block diagonal with 1x1 and 2x2 blocks, partial pivoting after
//                            Bunch and Kaufman: element growth bounded like the pivoted LU's)
// needs half the flops and -- decisive here -- half the shared memory of the general LU: the
// packed lower triangle of a k = 100 system is 42 KB, so five targets are resident per SM with
// ONE warp each and the factorisation runs without a single CTA barrier (the multi-warp LU spent
// 22 % of its stall samples at barriers and another 34 % waiting, at two CTAs per SM).
//
// Storage: rows 0..N-1 (N = m + 1: the m neighbours and the Lagrange row) and the right-hand
// side as row N of the SAME packed lower triangle, row i at i (i + 1) / 2.  The elimination of a
// pivot updates every row below it, so row N leaves the factorisation as z = D^-1 L^-1 P b and
// only the back substitution  L' x' = z  remains.  Lane l owns rows l, l + 32, ...: the column
// scans (pivot search), the multipliers and the rank-1 / rank-2 updates read and write a lane's
// own rows (consecutive rows of the packed triangle fall into distinct banks: the triangular
// numbers are a complete residue system modulo 16), the multipliers of the OTHER factor come
// from the stored column by broadcast loads.
//
// Any pivot below 1e-10 max|K| and any non-finite entry or solution sends the target to the
// pivoted LU (fail list, exactly like the Cholesky kernel), which owns the rank decision.

constexpr double LDL_ALPHA = 0.6403882032022076;     // (1 + sqrt 17) / 8

__device__ __forceinline__ int ldl_tri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ unsigned ldl_hi_abs(double v) { return (unsigned)__double2hiint(v) & 0x7fffffffu; }
__device__ __forceinline__ double ldl_from_hi(unsigned hi) { return __hiloint2double((int)hi, 0); }

// 1 / x for a normal x: MUFU.RCP64H + two Newton steps (as the LU's pivots)
__device__ __forceinline__ double ldl_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = __fma_rn(y, __fma_rn(-x, y, 1.0), y);
    return __fma_rn(y, __fma_rn(-x, y, 1.0), y);
}

// gamma(D_s, D_t) of the neighbour pairs, strictly lower triangle (predictor.py:573-575); the
// same expressions as the LU's assembly, so both solvers factor the same matrix.
template <int LUV>
__device__ __forceinline__ void ldl_assemble(double *A, const double2 *pxy, const double *pt, int m,
                                             const SolveParams &P, int lane, double &amax, bool &bad)
{
    constexpr int U = 4;
    const int npairs = (m * (m - 1)) >> 1;
#pragma unroll 1
    for (int e0 = lane; e0 < npairs; e0 += 32 * U) {
        unsigned w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = (e0 + 32 * u < npairs) ? __ldg(P.pairtab + e0 + 32 * u) : 0x00000100u;
        double g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int a = w[u] & 0xff, b = (w[u] >> 8) & 0xff;       // b < a
            const double2 pa = pxy[a], pb = pxy[b];
            const double dx = pa.x - pb.x, dy = pa.y - pb.y;
            const double lt = fabs(pt[a] - pt[b]);                   // utils.py:19-22
            if (LUV == LU_GENERIC) {
                const double h = sqrt(dx * dx + dy * dy);            // utils.py:40-44
                g[u] = gamma_st(P.M, h, lt);
            } else {
                const double d2 = __fma_rn(dy, dy, dx * dx) + 1e-290;    // sqrt(0) -> 1e-145, no 0 * inf
                const double h = sqrt_cubic_lu(d2);
                const double gs = __fma_rn(h, __fma_rn(-P.pcs, d2, P.pas), P.pbs);
                const double gt = __fma_rn(lt, __fma_rn(-P.pct, lt * lt, P.pat), P.pbt);
                const double pr = gs * gt;
                g[u] = (LUV == LU_PRODUCT_SUM_SPH) ? __fma_rn(P.M.k, pr, pr) : pr;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (e0 + 32 * u < npairs) {
                const int a = w[u] & 0xff, b = (w[u] >> 8) & 0xff;
                A[ldl_tri(a) + b] = g[u];
                amax = fmax(amax, fabs(g[u]));                       // fmax ignores NaN
                bad = bad || !(fabs(g[u]) < INFINITY);
            }
        }
    }
}

// Trailing update after the pivot block at columns j0 .. j0 + RANK - 1:
//      a(i, c) -= w1_i l1_c (+ w2_i l2_c),   j0 + RANK <= c <= i < n,
// w = the lane's own (unscaled) entries of the pivot columns, l = the stored multipliers of row c
// (broadcast loads).  CU columns per iteration: all loads, then all FMAs, then all stores.
template <int NG, int RANK, int CU>
__device__ __forceinline__ void ldl_update(double *A, const int (&rb)[NG], const double (&w1)[NG],
                                           const double (&w2)[NG], int j0, int n, int lane)
{
    int c = j0 + RANK;
    int tcj = ldl_tri(c) + j0;                       // &l1 of row c
#pragma unroll 1
    for (; c < n; c += CU) {
        double l1[CU], l2[CU];
        {
            int t = tcj;
            const int t0 = tcj;
#pragma unroll
            for (int u = 0; u < CU; ++u) {
                const int cc = c + u;
                const int tt = cc < n ? t : t0;      // rows past the matrix are never used
                l1[u] = A[tt];
                if (RANK == 2) l2[u] = A[tt + 1];
                t += cc + 1;                         // tri(cc + 1) = tri(cc) + cc + 1
            }
            tcj = t;
        }
        const int gfirst = c >> 5;                   // groups above hold no row >= c
        double a[NG][CU];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int i = lane + 32 * g;
            if (g >= gfirst && i < n) {
#pragma unroll
                for (int u = 0; u < CU; ++u)
                    if (i >= c + u) a[g][u] = A[rb[g] + c + u];
            }
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
#pragma unroll
            for (int u = 0; u < CU; ++u) {
                a[g][u] = __fma_rn(-w1[g], l1[u], a[g][u]);
                if (RANK == 2) a[g][u] = __fma_rn(-w2[g], l2[u], a[g][u]);
            }
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int i = lane + 32 * g;
            if (g >= gfirst && i < n) {
#pragma unroll
                for (int u = 0; u < CU; ++u)
                    if (i >= c + u) A[rb[g] + c + u] = a[g][u];
            }
        }
    }
}

// Symmetric interchange of rows / columns p < r of the packed lower triangle (all n rows: the
// finished columns of L and the right-hand-side row move with them).  Entry (r, p) stays.
template <int NG>
__device__ __forceinline__ void ldl_swap(double *A, const int (&rb)[NG], unsigned char *perm, int p, int r,
                                         int n, int lane)
{
    const int tp = ldl_tri(p), tr = ldl_tri(r);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const int i = lane + 32 * g;
        if (i < n && i != r) {
            int o1, o2;
            if (i < p) { o1 = tp + i; o2 = tr + i; }
            else if (i == p) { o1 = tp + p; o2 = tr + r; }
            else if (i < r) { o1 = rb[g] + p; o2 = tr + i; }
            else { o1 = rb[g] + p; o2 = rb[g] + r; }
            const double t1 = A[o1], t2 = A[o2];
            A[o1] = t2;
            A[o2] = t1;
        }
    }
    if (lane == 0) {
        const unsigned char t = perm[p];
        perm[p] = perm[r];
        perm[r] = t;
    }
    __syncwarp();
}

template <int NG, int LUV>
__global__ void __launch_bounds__(32) solve_ldl_kernel(SolveParams P, CholParams X, int adbl)
{
    extern __shared__ __align__(16) double smem_d[];
    constexpr int CU = 4;
    const int lane = threadIdx.x;
    const int k = P.k;
    const int kp = (k + 1) & ~1;
    double *A = smem_d;                                               // packed lower triangle, adbl doubles
    double2 *pxy = reinterpret_cast<double2 *>(A + adbl);
    double *pt = reinterpret_cast<double *>(pxy + kp);
    unsigned char *perm = reinterpret_cast<unsigned char *>(pxy);     // pxy is dead after the assembly
    double *sol = pt;                                                 // and so is pt
    // the epilogue's arrays alias the matrix, which is dead after the back substitution
    double *g0 = A, *rs = A + kp, *sc64 = A + 2 * kp;
    float *sc32a = reinterpret_cast<float *>(A + 3 * kp), *sc32b = sc32a + kp;
    unsigned *ids = reinterpret_cast<unsigned *>(A + 4 * kp);
    int rb[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) rb[g] = ldl_tri(lane + 32 * g);
    const double gdiag = gamma_st(P.M, 0.0, 0.0);                     // diagonal = gamma(0, 0), no special case

    // The next target's selection is fetched into registers while the current one is factorised.
    unsigned tl = 0, ntl = 0, nid[NG];
    int m = 0, nm = 0;
    double4 no[NG];
    double ng[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) {
        nid[i] = 0u;
        no[i] = make_double4(0.0, 0.0, 0.0, 0.0);
        ng[i] = 0.0;
    }
    bool has_next = false;
    auto fetch_ids = [&]() {
        nm = has_next ? P.sel_n[ntl] : 0;
#pragma unroll
        for (int i = 0; i < NG; ++i) {
            const int e = lane + 32 * i;
            if (has_next && e < k) nid[i] = P.sel_idx[(size_t)ntl * k + e];
        }
    };
    auto fetch_obs = [&]() {
#pragma unroll
        for (int i = 0; i < NG; ++i) {
            const int e = lane + 32 * i;
            if (e < nm) {
                no[i] = X.obs4[nid[i]];
                ng[i] = P.sel_gamma[(size_t)ntl * k + e];
            }
        }
    };
    if ((int)blockIdx.x < P.n) {
        has_next = true;
        ntl = P.order[blockIdx.x];
        fetch_ids();
        fetch_obs();
        tl = ntl;
        m = nm;
    }
    for (int slot = blockIdx.x; slot < P.n; slot += gridDim.x, tl = ntl, m = nm) {
        const int64_t tg = P.c0 + tl;
        const int N = m + 1, n = m + 2;
        const int tN = ldl_tri(N);
        __syncwarp();
        // ---- neighbours (prefetched), diagonal, Lagrange row, right-hand-side row ----
        unsigned cid[NG];
        double cres[NG], cg[NG];
        bool bad = !(fabs(gdiag) < INFINITY);
        if (m > 0) {
#pragma unroll
            for (int i = 0; i < NG; ++i) {
                const int e = lane + 32 * i;
                cid[i] = nid[i];
                cres[i] = no[i].w;
                cg[i] = ng[i];
                if (e < m) {
                    pxy[e] = make_double2(no[i].x, no[i].y);
                    pt[e] = no[i].z;
                    A[rb[i] + e] = gdiag;
                    A[ldl_tri(m) + e] = 1.0;          // Lagrange row            predictor.py:578
                    A[tN + e] = ng[i];                // rhs = [gamma_vector; 1]   predictor.py:582-585
                    bad = bad || !(fabs(ng[i]) < INFINITY);
                }
            }
            if (lane == 0) {
                A[ldl_tri(m) + m] = 0.0;              // predictor.py:580
                A[tN + m] = 1.0;
                A[tN + N] = 0.0;
            }
        }
        has_next = slot + (int)gridDim.x < P.n;
        if (has_next) ntl = P.order[slot + gridDim.x];
        if (m <= 0) {
            fetch_ids();
            fetch_obs();
            if (lane == 0) {
                P.mean[tg] = 0.0;
                P.std[tg] = 0.0;
                if (P.nused_out) P.nused_out[tg] = m < 0 ? -1 : 0;
            }
            for (int e = lane; e < k; e += 32) {
                if (P.w_out) P.w_out[(size_t)tg * k + e] = 0.f;
                if (P.g_out) P.g_out[(size_t)tg * k + e] = 0.f;
                if (P.idx_out) P.idx_out[(size_t)tg * k + e] = 0u;
            }
            continue;
        }
        __syncwarp();
        double amax = fmax(fabs(gdiag), 1.0);         // the border holds ones
        if (m > 1) ldl_assemble<LUV>(A, pxy, pt, m, P, lane, amax, bad);
        amax = __longlong_as_double((long long)warp_max_u64((unsigned long long)__double_as_longlong(amax)));
        const double tol = 1e-10 * amax;
        bool fail = __any_sync(FULL, bad);
        __syncwarp();                                 // pxy is dead from here: perm lives there
#pragma unroll
        for (int g = 0; g < NG; ++g)
            if (lane + 32 * g < N) perm[lane + 32 * g] = (unsigned char)(lane + 32 * g);
        __syncwarp();
        fetch_ids();

        // ---- Bunch-Kaufman: pivots among rows / columns 0 .. N-1 ----
        int j = 0;
        while (j < N && !fail) {
            const int tj = ldl_tri(j);
            double v[NG];
            unsigned key = 0u;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int i = lane + 32 * g;
                v[g] = 0.0;
                if (i > j && i < n) {
                    v[g] = A[rb[g] + j];
                    if (i < N) key = max(key, (ldl_hi_abs(v[g]) & ~0xffu) | (unsigned)i);
                }
            }
            double d = A[tj + j];
            key = __reduce_max_sync(FULL, key);
            const int r = (int)(key & 0xffu);
            const double lam = ldl_from_hi(key & ~0xffu);      // max |a(i, j)|, i > j (truncated to 12 bits)
            const double aj = fabs(d);
            int kind = 0;                                      // 0: 1x1; 1: 1x1 after j <-> r; 2: 2x2 with r
            if (!(aj >= LDL_ALPHA * lam)) {
                if (r <= j) { fail = true; break; }            // NaN pivot
                const int tr = ldl_tri(r);
                unsigned sk = 0u;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int c = lane + 32 * g;
                    if (c >= j && c < N && c != r) {
                        const double val = (c < r) ? A[tr + c] : A[rb[g] + r];
                        sk = max(sk, ldl_hi_abs(val));
                    }
                }
                const double arr = fabs(A[tr + r]);
                sk = __reduce_max_sync(FULL, sk);
                const double sig = ldl_from_hi(sk);            // max |a(r, c)|, c != r
                if (aj * sig >= LDL_ALPHA * lam * lam) kind = 0;
                else if (arr >= LDL_ALPHA * sig) kind = 1;
                else kind = 2;
            }
            if (kind == 1) {
                ldl_swap<NG>(A, rb, perm, j, r, n, lane);
                d = A[tj + j];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int i = lane + 32 * g;
                    v[g] = (i > j && i < n) ? A[rb[g] + j] : 0.0;
                }
            } else if (kind == 2 && r != j + 1) {
                ldl_swap<NG>(A, rb, perm, j + 1, r, n, lane);
            }
            if (kind < 2) {
                if (!(fabs(d) > tol)) { fail = true; break; }
                const double dinv = ldl_rcp(d);
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int i = lane + 32 * g;
                    if (i > j && i < n) A[rb[g] + j] = v[g] * dinv;
                }
                __syncwarp();
                ldl_update<NG, 1, CU>(A, rb, v, v, j, n, lane);
                __syncwarp();
                j += 1;
            } else {
                if (!(lam > tol)) { fail = true; break; }
                const int tj1 = tj + j + 1;
                const double d11 = A[tj + j], d21 = A[tj1 + j], d22 = A[tj1 + j + 1];
                // [l1 l2] = [w1 w2] D^-1, scaled by the off-diagonal like LAPACK's dsytf2
                const double t = ldl_rcp(d21);
                const double ak = d11 * t, akp1 = d22 * t;
                const double den = t * ldl_rcp(__fma_rn(ak, akp1, -1.0));
                double w1[NG], w2[NG];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int i = lane + 32 * g;
                    w1[g] = 0.0;
                    w2[g] = 0.0;
                    if (i > j + 1 && i < n) {
                        w1[g] = A[rb[g] + j];
                        w2[g] = A[rb[g] + j + 1];
                        A[rb[g] + j] = den * __fma_rn(akp1, w1[g], -w2[g]);
                        A[rb[g] + j + 1] = den * __fma_rn(ak, w2[g], -w1[g]);
                    }
                }
                __syncwarp();
                if (lane == 0) A[tj1 + j] = 0.0;               // L(j+1, j) = 0: D is not needed again
                ldl_update<NG, 2, CU>(A, rb, w1, w2, j, n, lane);
                __syncwarp();
                j += 2;
            }
        }
        fetch_obs();

        // ---- back substitution L' x' = z (row N), row by row: x'_i is final once the rows below it
        //      are done; lane l keeps z of columns l, l + 32, ... ----
        double zr[NG];
        bool finite = true;
        if (!fail) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int c = lane + 32 * g;
                zr[g] = (c < N) ? A[tN + c] : 0.0;
            }
#pragma unroll
            for (int gi = NG - 1; gi >= 0; --gi) {
                if (32 * gi >= N) continue;
                int i = min(32 * gi + 31, N - 1);
                int ti = ldl_tri(i);
#pragma unroll 1
                for (; i >= 32 * gi && i >= 1; --i) {
                    const double xi = __shfl_sync(FULL, zr[gi], i & 31);
#pragma unroll
                    for (int g = 0; g <= gi; ++g) {
                        const int c = lane + 32 * g;
                        if (c < i) zr[g] = __fma_rn(-A[ti + c], xi, zr[g]);
                    }
                    ti -= i;                                   // tri(i - 1) = tri(i) - i
                }
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int c = lane + 32 * g;
                if (c < N) {
                    finite = finite && (fabs(zr[g]) < INFINITY);
                    const int o = perm[c];
                    if (o < m) sol[o] = zr[g];                 // o == m: the Lagrange multiplier
                }
            }
            fail = !__all_sync(FULL, finite);
        }
        if (fail) {   // tiny pivot / non-finite: the pivoted LU takes this target (and decides its rank)
            if (lane == 0) X.fail_list[atomicAdd(X.fail_count, 1u)] = tl;
            continue;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NG; ++i) {
            const int e = lane + 32 * i;
            if (e < m) {
                g0[e] = cg[i];
                rs[e] = cres[i];
                ids[e] = cid[i];
            }
        }
        __syncwarp();
        solve_epilogue(P, sol, g0, ids, rs, sc64, sc32a, sc32b, m, tg, lane);
    }
}
