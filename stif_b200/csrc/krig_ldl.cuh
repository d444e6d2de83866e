// krig_ldl.cuh -- symmetric indefinite kriging solve: blocked Bunch-Kaufman LDL' of the reference's
// own bordered system, trailing updates on the FP64 tensor cores; one warp per target up to
// k = 62, one warp per 32-row group above.  Included by kriging.cu (namespace stif::krig).
//
// Replaces calc_kriging_weights (stif/predictor.py:566-593) for models that are NOT valid
// variograms: product / product_sum (stif/variogram_models.py:126-127, :155-157) give an
// indefinite gamma matrix, unconstrained Nelder-Mead fits a negative total nugget.  The reference
// solves  K x = b,  K = [Gamma 1; 1' 0],  b = [gamma0; 1]  with LAPACK dgelsd; K is symmetric, so
//      P K P' = L D L'      (D block diagonal with 1x1 and 2x2 blocks, partial pivoting after
//                            Bunch and Kaufman: element growth bounded like the pivoted LU's)
// needs half the flops and -- decisive here -- half the shared memory of the general LU: the
// packed lower triangle of a k = 100 system is 42 KB against 92 KB, five targets are resident per
// SM instead of two.
//
// Storage: rows 0..N-1 (N = m + 1: the m neighbours and the Lagrange row) and the right-hand
// side as row N of the SAME packed lower triangle, row i at i (i + 1) / 2.  The elimination of a
// pivot updates every row below it, so row N leaves the factorisation as z = D^-1 L^-1 P b and
// only the back substitution  L' x' = z  remains.  Thread (warp, lane) owns rows
// lane + 32 (warp + W s) (consecutive rows of the packed triangle fall into distinct banks: the
// triangular numbers are a complete residue system modulo 16).
//
// Blocking (the structure of LAPACK's dsytrf / dlasyf, without its workspace): inside a panel of
// up to eight columns the matrix is touched by the symmetric interchanges only.  The column under
// examination -- and row r of the Bunch-Kaufman test, by symmetry -- is brought up to date on the
// fly,
//      a(i, j) - sum_q L(i, q) W(j, q),     W = L D  recomputed from D (16 doubles of shared memory),
// and the multipliers L(:, j) are stored in place.  After the panel the whole trailing matrix
// receives  - L W'  in 8x8 tiles of two mma.sync.m8n8k4.f64 (DMMA) each, masked at the edges of
// the triangle.
//
// Warps per target: with one warp a target needs no CTA barrier at all, and up to k = 62 sixteen
// such targets are resident per SM.  At k = 100 only five fit, too few to hide the latency of the
// dependent panel steps, so every 32-row group gets its own warp (W = NG): the warps take identical
// decisions from shared data (pivot keys meet in a double-buffered slot + one barrier), tile
// columns of the trailing update are dealt round-robin, warp 0 does the back substitution and the
// epilogue.  Results are bit-identical for every W (measured at C5, 125k targets: 25.9 ms with one
// warp, 21.0 with two, 19.1 with four; the pivoted LU: 34.7).
//
// Any pivot below 1e-10 max|K| and any non-finite entry or solution sends the target to the
// pivoted LU (fail list, exactly like the Cholesky kernel), which owns the rank decision.

constexpr double LDL_ALPHA = 0.6403882032022076;     // (1 + sqrt 17) / 8
#ifndef STIF_LDL_PW
#define STIF_LDL_PW 8
#endif
constexpr int LDL_PW = STIF_LDL_PW;                  // panel width (<= 8: two DMMA k-steps of 4)

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
template <int LUV, int NT>
__device__ __forceinline__ void ldl_assemble(double *A, const double2 *pxy, const double *pt, int m,
                                             const SolveParams &P, int tid, unsigned &amax_hi)
{
    constexpr int U = 4;
    const int npairs = (m * (m - 1)) >> 1;
#pragma unroll 1
    for (int e0 = tid; e0 < npairs; e0 += NT * U) {
        unsigned w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = (e0 + NT * u < npairs) ? __ldg(P.pairtab + e0 + NT * u) : 0x00000100u;
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
            if (e0 + NT * u < npairs) {
                const int a = w[u] & 0xff, b = (w[u] >> 8) & 0xff;
                A[ldl_tri(a) + b] = g[u];
                amax_hi = max(amax_hi, ldl_hi_abs(g[u]));            // Inf / NaN: >= 0x7ff00000
            }
        }
    }
}

// Trailing update after a panel at columns j0 .. j0 + pw - 1 (pw <= 8; j1 = j0 + pw):
//      a(i, c) -= sum_q L(i, j0 + q) W(c, q),    j1 <= c <= i < n,    W = L D,
// in 8x8 tiles of the absolute 8-grid: two DMMA per tile (A fragment = multipliers of the tile's
// rows, B fragment = -W of the tile's columns, built from the multipliers and this lane's entries of
// D), C fragment loaded / stored under the mask of the triangle and of the finished columns.  Rows
// and columns outside the mask only pollute their own (masked) outputs.  Two tiles per iteration.
// Two tiles (rows 8 I + g and 8 I + 8 + g, tile column J).  EDGE: the mask of the triangle, of the
// finished columns and of the matrix end is applied; interior tile pairs need none of it.
template <bool EDGE>
__device__ __forceinline__ void ldl_tile_pair(double *A, int rowx, int tx, int cj, int colA, int j1, int n,
                                              bool secondrow, bool k0, bool k1, double b0, double b1)
{
    const int rowy = rowx + 8;
    const int ty = tx + 8 * rowx + 36;                   // tri(r + 8) = tri(r) + 8 r + 36
    if (EDGE) {
        const bool inx = rowx < n, iny = rowy < n && secondrow;
        const bool rokx = inx && rowx >= j1, roky = iny && rowy >= j1;
        const bool cokA = colA >= j1, cokB = colA + 1 >= j1;
        const bool xA = rokx && cokA && colA <= rowx, xB = rokx && cokB && colA + 1 <= rowx;
        const bool yA = roky && cokA && colA <= rowy, yB = roky && cokB && colA + 1 <= rowy;
        const double ax0 = (k0 && inx) ? A[tx + cj] : 0.0, ax1 = (k1 && inx) ? A[tx + cj + 4] : 0.0;
        const double ay0 = (k0 && iny) ? A[ty + cj] : 0.0, ay1 = (k1 && iny) ? A[ty + cj + 4] : 0.0;
        double cx0 = xA ? A[tx + colA] : 0.0, cx1 = xB ? A[tx + colA + 1] : 0.0;
        double cy0 = yA ? A[ty + colA] : 0.0, cy1 = yB ? A[ty + colA + 1] : 0.0;
        dmma884(cx0, cx1, ax0, b0);
        dmma884(cy0, cy1, ay0, b0);
        dmma884(cx0, cx1, ax1, b1);
        dmma884(cy0, cy1, ay1, b1);
        if (xA) A[tx + colA] = cx0;
        if (xB) A[tx + colA + 1] = cx1;
        if (yA) A[ty + colA] = cy0;
        if (yB) A[ty + colA + 1] = cy1;
    } else {
        const double ax0 = A[tx + cj], ax1 = k1 ? A[tx + cj + 4] : 0.0;
        const double ay0 = A[ty + cj], ay1 = k1 ? A[ty + cj + 4] : 0.0;
        double cx0 = A[tx + colA], cx1 = A[tx + colA + 1];
        double cy0 = A[ty + colA], cy1 = A[ty + colA + 1];
        dmma884(cx0, cx1, ax0, b0);
        dmma884(cy0, cy1, ay0, b0);
        dmma884(cx0, cx1, ax1, b1);
        dmma884(cy0, cy1, ay1, b1);
        A[tx + colA] = cx0;
        A[tx + colA + 1] = cx1;
        A[ty + colA] = cy0;
        A[ty + colA + 1] = cy1;
    }
}

template <int W>
__device__ __forceinline__ void ldl_trailing(double *A, const double *dpan, int j0, int pw, int j1, int n,
                                             bool has22, int lane, int warp)
{
    const int g = lane >> 2, t = lane & 3;
    // this lane's entries of D for the panel columns t and 4 + t: diagonal, coupling with the next
    // and with the previous column (a block never starts at the panel's last column)
    const double d0 = dpan[2 * t], d1 = dpan[2 * (4 + t)];
    const double u0 = dpan[2 * t + 1], u1 = (t < 3) ? dpan[2 * (4 + t) + 1] : 0.0;
    const double l0 = (t > 0) ? dpan[2 * (t - 1) + 1] : 0.0, l1 = dpan[2 * (3 + t) + 1];
    const bool k0 = t < pw, k1 = 4 + t < pw;
    const int nT = (n + 7) >> 3;
    const int cj = j0 + t;
    const int Jf = j1 >> 3;
    const bool k0all = pw >= 4;                          // every lane's first k-column is a finished one
    // tile columns dealt to the warps, every other round in reverse order: column J holds nT - J tiles
#pragma unroll 1
    for (int it = 0;; ++it) {
        const int J = Jf + it * W + ((it & 1) ? W - 1 - warp : warp);
        if (J >= nT) break;
        double b0 = 0.0, b1 = 0.0;
        {
            const int c = min(8 * J + g, n - 1);
            const int tc = ldl_tri(c) + cj;
            if (k0) {
                b0 = d0 * A[tc];
                if (has22) {
                    if (u0 != 0.0) b0 = __fma_rn(u0, A[tc + 1], b0);
                    if (l0 != 0.0) b0 = __fma_rn(l0, A[tc - 1], b0);
                }
            }
            if (k1) {
                b1 = d1 * A[tc + 4];
                if (has22) {
                    if (u1 != 0.0) b1 = __fma_rn(u1, A[tc + 5], b1);
                    if (l1 != 0.0) b1 = __fma_rn(l1, A[tc + 3], b1);
                }
            }
            b0 = -b0;
            b1 = -b1;
        }
        const int colA = 8 * J + 2 * t;
        int rowx = 8 * J + g;
        int tx = ldl_tri(rowx);
        // the diagonal tile (and its lower neighbour) under the mask
        ldl_tile_pair<true>(A, rowx, tx, cj, colA, j1, n, J + 1 < nT, k0, k1, b0, b1);
        tx += 16 * rowx + 136;                           // tri(r + 16) = tri(r) + 16 r + 136
        rowx += 16;
        int I = J + 2;
        if (J > Jf && k0all) {                           // interior: all columns >= j1, all rows below the diagonal
            for (; 8 * I + 15 < n; I += 2) {
                ldl_tile_pair<false>(A, rowx, tx, cj, colA, j1, n, true, true, k1, b0, b1);
                tx += 16 * rowx + 136;
                rowx += 16;
            }
        }
#pragma unroll 1
        for (; I < nT; I += 2) {
            ldl_tile_pair<true>(A, rowx, tx, cj, colA, j1, n, I + 1 < nT, k0, k1, b0, b1);
            tx += 16 * rowx + 136;
            rowx += 16;
        }
    }
}

// Row ownership: W warps per target; thread (warp, lane) owns rows r0 + RS s, r0 = lane + 32 warp,
// RS = 32 W, s = 0 .. GW-1 (rows >= n are idle).  W = 1: one warp, barriers are __syncwarp().
template <int W>
__device__ __forceinline__ void ldl_sync()
{
    if constexpr (W == 1) __syncwarp();
    else __syncthreads();
}

// max over the whole target (all W warps); slots: 2 W words, double-buffered by `phase`
template <int W>
__device__ __forceinline__ unsigned ldl_allmax(unsigned v, unsigned *slots, int &phase, int warp, int lane)
{
    v = __reduce_max_sync(FULL, v);
    if constexpr (W > 1) {
        if (lane == 0) slots[phase * W + warp] = v;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < W; ++w) v = max(v, slots[phase * W + w]);
        phase ^= 1;
    }
    return v;
}

// Symmetric interchange of rows / columns p < r of the packed lower triangle (all n rows: the
// finished columns of L and the right-hand-side row move with them).  Entry (r, p) stays.
template <int GW, int RS, int W>
__device__ __forceinline__ void ldl_swap(double *A, const int (&rb)[GW], unsigned char *perm, int p, int r,
                                         int n, int r0, int tid)
{
    const int tp = ldl_tri(p), tr = ldl_tri(r);
#pragma unroll
    for (int s = 0; s < GW; ++s) {
        const int i = r0 + RS * s;
        if (i < n && i != r) {
            int o1, o2;
            if (i < p) { o1 = tp + i; o2 = tr + i; }
            else if (i == p) { o1 = tp + p; o2 = tr + r; }
            else if (i < r) { o1 = rb[s] + p; o2 = tr + i; }
            else { o1 = rb[s] + p; o2 = rb[s] + r; }
            const double t1 = A[o1], t2 = A[o2];
            A[o1] = t2;
            A[o2] = t1;
        }
    }
    if (tid == 0) {
        const unsigned char t = perm[p];
        perm[p] = perm[r];
        perm[r] = t;
    }
    ldl_sync<W>();
}

// D of the current panel lives in shared memory as pairs: dpan[2 q] = the diagonal entry of panel column
// q, dpan[2 q + 1] = the coupling of a 2x2 block that starts at q (bit q of the uniform mask m22), so
// one 16-byte broadcast load fetches both.
//
// Column x >= j (the column under examination, row r of the Bunch-Kaufman test by symmetry, or the
// second column of a 2x2 block) brought up to date over the q finished panel columns,
//      a(c, x) - sum_qq L(c, qq) W(x, qq),     W(x, qq) = (L D)(x, qq)  (broadcast loads),
// for the thread's rows `on` and for the diagonal entry.  The loop over the (at most eight) finished
// columns is unrolled with a uniform early exit, so every address is base + constant.
template <int GW, int RS, bool BELOW>        // BELOW: every row in `on` lies below x (no symmetric access)
__device__ __forceinline__ void ldl_col_generic(const double *A, const double *dpan, const int (&rb)[GW], int j0,
                                                int q, int x, unsigned m22, const bool (&on)[GW], int r0,
                                                double (&val)[GW], double &diag)
{
    const int tx = ldl_tri(x);
    const double *lx = A + tx + j0;          // row x's multipliers
    const double *lp[GW];                    // the thread's own rows' multipliers
#pragma unroll
    for (int s = 0; s < GW; ++s) {
        const int c = r0 + RS * s;
        lp[s] = A + rb[s] + j0;
        const int o = (BELOW || c > x) ? rb[s] + x : tx + c;
        val[s] = on[s] ? A[o] : 0.0;
    }
    double dg = A[tx + x];
    if (m22 == 0u) {                         // no 2x2 block among the finished columns: W = d L
#pragma unroll
        for (int qq = 0; qq < LDL_PW; ++qq) {
            if (qq >= q) break;
            const double lxq = lx[qq];
            const double w = dpan[2 * qq] * lxq;
            dg = __fma_rn(-lxq, w, dg);
#pragma unroll
            for (int s = 0; s < GW; ++s)
                if (on[s]) val[s] = __fma_rn(-lp[s][qq], w, val[s]);
        }
    } else {
        // W(x, qq) = d L(x, qq) + off(qq) L(x, qq + 1) + off(qq - 1) L(x, qq - 1); off is zero wherever no
        // block couples the two columns, and L(x, qq + 1) is a finite entry of row x even when that
        // column is not finished yet (j0 + qq + 1 <= j <= x), so the products need no predicate
        double lprev = 0.0, offprev = 0.0, lcur = lx[0];
#pragma unroll
        for (int qq = 0; qq < LDL_PW; ++qq) {
            if (qq >= q) break;
            const double2 dq = *reinterpret_cast<const double2 *>(dpan + 2 * qq);     // {d, off}
            const double lnext = lx[qq + 1], offq = dq.y;
            double w = dq.x * lcur;
            w = __fma_rn(offq, lnext, w);
            w = __fma_rn(offprev, lprev, w);
            dg = __fma_rn(-lcur, w, dg);
#pragma unroll
            for (int s = 0; s < GW; ++s)
                if (on[s]) val[s] = __fma_rn(-lp[s][qq], w, val[s]);
            lprev = lcur;
            lcur = lnext;
            offprev = offq;
        }
    }
    diag = dg;
}

// Back substitution (warp 0), rows of group GI (rows 32 GI .. 32 GI + 31) from the bottom up: x'_i is
// final once the rows below it are done; lane l keeps z of columns l, l + 32, ...  Template recursion
// keeps every register index a compile-time constant.
template <int NG, int GI>
__device__ __forceinline__ void ldl_backsub(const double *A, double (&zr)[NG], int N, int lane)
{
    if constexpr (GI >= 0) {
        if (32 * GI < N) {
            int i = min(32 * GI + 31, N - 1);
            int ti = ldl_tri(i);
#pragma unroll 1
            for (; i >= 32 * GI && i >= 1; --i) {
                const double xi = __shfl_sync(FULL, zr[GI], i & 31);
#pragma unroll
                for (int g = 0; g <= GI; ++g) {
                    const int c = lane + 32 * g;
                    if (c < i) zr[g] = __fma_rn(-A[ti + c], xi, zr[g]);
                }
                ti -= i;                                   // tri(i - 1) = tri(i) - i
            }
        }
        ldl_backsub<NG, GI - 1>(A, zr, N, lane);
    }
}

// NG: 32-row groups of the k + 2 rows; W: warps per target (row groups dealt round-robin).
template <int NG, int W, int LUV>
__global__ void __launch_bounds__(32 * W) solve_ldl_kernel(SolveParams P, CholParams X, int adbl)
{
    extern __shared__ __align__(16) double smem_d[];
    constexpr int GW = (NG + W - 1) / W;      // row slots per thread
    constexpr int RS = 32 * W;                // row stride between a thread's slots
    constexpr int NT = 32 * W;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = tid;                       // = lane + 32 warp
    const int k = P.k;
    const int kp = (k + 1) & ~1;
    double *A = smem_d;                                               // packed lower triangle, adbl doubles
    double2 *pxy = reinterpret_cast<double2 *>(A + adbl);
    double *pt = reinterpret_cast<double *>(pxy + kp);
    unsigned char *perm = reinterpret_cast<unsigned char *>(pxy);     // pxy is dead after the assembly
    double *sol = pt;                                                 // and so is pt
    double *dpan = pt + kp;                                           // D of the current panel: 8 x {diagonal, coupling}, 16-byte aligned
    double *sxch = dpan + 16;                                         // three doubles handed between rows' owners
    unsigned *slots = reinterpret_cast<unsigned *>(sxch + 3);         // [2][W] max exchange (<= 10 words)
    // the epilogue's arrays alias the matrix, which is dead after the back substitution
    double *g0 = A, *rs = A + kp, *sc64 = A + 2 * kp;
    float *sc32a = reinterpret_cast<float *>(A + 3 * kp), *sc32b = sc32a + kp;
    unsigned *ids = reinterpret_cast<unsigned *>(A + 4 * kp);
    int rb[GW];
#pragma unroll
    for (int s = 0; s < GW; ++s) rb[s] = ldl_tri(r0 + RS * s);
    const double gdiag = gamma_st(P.M, 0.0, 0.0);                     // diagonal = gamma(0, 0), no special case
    int phase = 0;

    // The next target's selection is fetched into registers while the current one is factorised
    // (thread -> neighbours r0 + RS s, the same ownership as the rows).
    unsigned tl = 0, ntl = 0, nid[GW];
    int m = 0, nm = 0;
    double4 no[GW];
    double ng[GW];
#pragma unroll
    for (int i = 0; i < GW; ++i) {
        nid[i] = 0u;
        no[i] = make_double4(0.0, 0.0, 0.0, 0.0);
        ng[i] = 0.0;
    }
    bool has_next = false;
    auto fetch_ids = [&]() {
        nm = has_next ? P.sel_n[ntl] : 0;
#pragma unroll
        for (int i = 0; i < GW; ++i) {
            const int e = r0 + RS * i;
            if (has_next && e < k) nid[i] = P.sel_idx[(size_t)ntl * k + e];
        }
    };
    auto fetch_obs = [&]() {
#pragma unroll
        for (int i = 0; i < GW; ++i) {
            const int e = r0 + RS * i;
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
        ldl_sync<W>();                                // the previous target's epilogue is done with the arrays
        // ---- neighbours (prefetched), diagonal, Lagrange row, right-hand-side row ----
        unsigned cid[GW];
        double cres[GW], cg[GW];
        bool bad = !(fabs(gdiag) < INFINITY);
        if (m > 0) {
#pragma unroll
            for (int i = 0; i < GW; ++i) {
                const int e = r0 + RS * i;
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
            if (tid == 0) {
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
            if (tid == 0) {
                P.mean[tg] = 0.0;
                P.std[tg] = 0.0;
                if (P.nused_out) P.nused_out[tg] = m < 0 ? -1 : 0;
            }
            for (int e = tid; e < k; e += NT) {
                if (P.w_out) P.w_out[(size_t)tg * k + e] = 0.f;
                if (P.g_out) P.g_out[(size_t)tg * k + e] = 0.f;
                if (P.idx_out) P.idx_out[(size_t)tg * k + e] = 0u;
            }
            continue;
        }
        ldl_sync<W>();
        // largest |entry| (high word: 2^-20 resolution is plenty for a tolerance); the border holds ones;
        // NaN / Inf among the entries show as a high word >= 0x7ff00000
        unsigned amax_hi = max(ldl_hi_abs(gdiag), 0x3ff00000u);
        if (bad) amax_hi = 0x7ff00000u;
        if (m > 1) ldl_assemble<LUV, NT>(A, pxy, pt, m, P, tid, amax_hi);
        amax_hi = ldl_allmax<W>(amax_hi, slots, phase, warp, lane);
        const double tol = 1e-10 * ldl_from_hi(amax_hi);
        bool fail = amax_hi >= 0x7ff00000u;           // uniform over the target's warps
        ldl_sync<W>();                                // pxy is dead from here: perm lives there
#pragma unroll
        for (int s = 0; s < GW; ++s)
            if (r0 + RS * s < N) perm[r0 + RS * s] = (unsigned char)(r0 + RS * s);
        fetch_ids();

        // ---- blocked Bunch-Kaufman: pivots among rows / columns 0 .. N-1 (uniform control flow) ----
        int j0 = 0;
        while (j0 < N && !fail) {
            if (tid < 16) dpan[tid] = 0.0;
            ldl_sync<W>();
            unsigned m22 = 0u;                  // bit q: a 2x2 block starts at panel column q
            int q = 0;                          // finished columns of this panel
            while (q < LDL_PW && j0 + q < N) {
                const int j = j0 + q, tj = ldl_tri(j);
                bool act[GW];
#pragma unroll
                for (int s = 0; s < GW; ++s) act[s] = r0 + RS * s > j && r0 + RS * s < n;
                double v[GW], d;
                ldl_col_generic<GW, RS, true>(A, dpan, rb, j0, q, j, m22, act, r0, v, d);
                unsigned key = 0u;
#pragma unroll
                for (int s = 0; s < GW; ++s) {
                    const int i = r0 + RS * s;
                    if (act[s] && i < N) key = max(key, (ldl_hi_abs(v[s]) & ~0xffu) | (unsigned)i);
                }
                key = ldl_allmax<W>(key, slots, phase, warp, lane);
                const int r = (int)(key & 0xffu);
                const double lam = ldl_from_hi(key & ~0xffu);  // max |a(i, j)|, i > j (truncated to 12 bits)
                const double aj = fabs(d);
                int kind = 0;                   // 0 / 1: 1x1 pivot (1: after j <-> r); 2: 2x2 pivot (j, r)
                double val[GW], arr = 0.0;      // column r up to date (row r of the test, by symmetry)
#pragma unroll
                for (int s = 0; s < GW; ++s) val[s] = 0.0;
                if (!(aj >= LDL_ALPHA * lam)) {
                    if (r <= j) { fail = true; break; }            // NaN pivot
                    // sigma = max |a(r, c)| over c != r (c = j: lam); the values are kept: after an
                    // interchange column r IS the pivot column and needs no second evaluation
                    bool on[GW];
#pragma unroll
                    for (int s = 0; s < GW; ++s) on[s] = act[s] && r0 + RS * s != r;
                    ldl_col_generic<GW, RS, false>(A, dpan, rb, j0, q, r, m22, on, r0, val, arr);
                    unsigned sk = key & ~0xffu;
#pragma unroll
                    for (int s = 0; s < GW; ++s)
                        if (on[s] && r0 + RS * s < N) sk = max(sk, ldl_hi_abs(val[s]));
                    sk = ldl_allmax<W>(sk, slots, phase, warp, lane);
                    const double sig = ldl_from_hi(sk);
                    if (aj * sig >= LDL_ALPHA * lam * lam) kind = 0;
                    else if (fabs(arr) >= LDL_ALPHA * sig) kind = 1;
                    else kind = 2;
                }
                if (kind == 2 && q == LDL_PW - 1) break;   // the 2x2 block would leave the panel: close it
                if (kind == 1) {
                    // rows / columns j <-> r: the new column j is the old column r (row r keeps a(r, j))
                    ldl_sync<W>();                         // every warp has read rows j and r
                    ldl_swap<GW, RS, W>(A, rb, perm, j, r, n, r0, tid);
#pragma unroll
                    for (int s = 0; s < GW; ++s)
                        if (r0 + RS * s != r) v[s] = val[s];
                    d = arr;
                }
                if (kind < 2) {
                    if (!(fabs(d) > tol)) { fail = true; break; }
                    const double dinv = ldl_rcp(d);
#pragma unroll
                    for (int s = 0; s < GW; ++s)
                        if (act[s]) A[rb[s] + j] = v[s] * dinv;
                    if (tid == 0) dpan[2 * q] = d;
                    q += 1;
                } else {
                    if (!(lam > tol)) { fail = true; break; }
                    // 2x2 pivot: rows / columns j + 1 <-> r, the second pivot column is the old column r.
                    // Three entries change owner: a(j+1, j) and a(j+1, r) move to row r, a(r, j) becomes d21.
                    const int pr = j + 1, tj1 = tj + j + 1;
#pragma unroll
                    for (int s = 0; s < GW; ++s) {
                        const int i = r0 + RS * s;
                        if (i == pr) {
                            sxch[0] = v[s];
                            sxch[1] = val[s];
                        }
                        if (i == r) sxch[2] = v[s];
                    }
                    ldl_sync<W>();                         // (also: every warp has read rows j + 1 and r)
                    const double xa = sxch[0], xb = sxch[1], d21 = sxch[2], d22 = arr;
                    if (r != pr) ldl_swap<GW, RS, W>(A, rb, perm, pr, r, n, r0, tid);
                    // [l1 l2] = [w1 w2] D^-1, scaled by the off-diagonal like LAPACK's dsytf2
                    const double t = ldl_rcp(d21);
                    const double ak = d * t, akp1 = d22 * t;
                    const double den = t * ldl_rcp(__fma_rn(ak, akp1, -1.0));
#pragma unroll
                    for (int s = 0; s < GW; ++s) {
                        const int i = r0 + RS * s;
                        if (act[s] && i > pr) {
                            const double w1 = (i == r) ? xa : v[s], w2 = (i == r) ? xb : val[s];
                            A[rb[s] + j] = den * __fma_rn(akp1, w1, -w2);
                            A[rb[s] + j + 1] = den * __fma_rn(ak, w2, -w1);
                        }
                    }
                    if (tid == 0) {
                        A[tj1 + j] = 0.0;                  // L(j+1, j) = 0
                        dpan[2 * q] = d;
                        dpan[2 * q + 1] = d21;
                        dpan[2 * q + 2] = d22;
                    }
                    m22 |= 1u << q;
                    q += 2;
                }
                ldl_sync<W>();
            }
            if (fail) break;
            if (j0 + q < N) ldl_trailing<W>(A, dpan, j0, q, j0 + q, n, m22 != 0u, lane, warp);
            ldl_sync<W>();
            j0 += q;
        }
        fetch_obs();

        // ---- back substitution L' x' = z (row N) by warp 0: lane l keeps z of columns l, l + 32, ... ----
        ldl_sync<W>();
        bool solved = !fail;
        if (warp == 0 && !fail) {
            double zr[NG];
            bool finite = true;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int c = lane + 32 * g;
                zr[g] = (c < N) ? A[tN + c] : 0.0;
            }
            ldl_backsub<NG, NG - 1>(A, zr, N, lane);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int c = lane + 32 * g;
                if (c < N) {
                    finite = finite && (fabs(zr[g]) < INFINITY);
                    const int o = perm[c];
                    if (o < m) sol[o] = zr[g];                 // o == m: the Lagrange multiplier
                }
            }
            solved = __all_sync(FULL, finite);
        }
        ldl_sync<W>();                                // the matrix is dead: stage the epilogue's inputs there
#pragma unroll
        for (int i = 0; i < GW; ++i) {
            const int e = r0 + RS * i;
            if (e < m) {
                g0[e] = cg[i];
                rs[e] = cres[i];
                ids[e] = cid[i];
            }
        }
        ldl_sync<W>();
        if (warp == 0) {
            if (solved) solve_epilogue(P, sol, g0, ids, rs, sc64, sc32a, sc32b, m, tg, lane);
            // tiny pivot / non-finite: the pivoted LU takes this target (and decides its rank)
            else if (lane == 0) X.fail_list[atomicAdd(X.fail_count, 1u)] = tl;
        }
    }
}
