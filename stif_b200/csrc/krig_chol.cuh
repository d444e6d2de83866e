// krig_chol.cuh -- covariance-form kriging solve: one warp per target, tiled Cholesky with the
// rank-4 updates on the FP64 tensor cores.  Included by kriging.cu (namespace stif::krig).
//
// Replaces calc_kriging_weights (stif/predictor.py:566-593) for VALID variogram models.
// The reference solves the bordered system  [Gamma 1; 1' 0] [w; lam] = [gamma0; 1]  with LAPACK
// dgelsd.  With sum(w) = 1 the same weights solve the covariance form
//      C w - lam 1 = c0,   C = c 11' - Gamma,  c0 = c 1 - gamma0,   c = gamma(inf, inf) (the sill),
// and C is symmetric positive definite whenever gamma is a valid variogram (sum / metric /
// sum_metric of spherical or gaussian marginals with non-negative sills and nuggets).  Then
//      C = L L',  yv = L^-1 c0,  yu = L^-1 1,  lam = (1 - yu.yv) / (yu.yu),  w = L^-T (yv + lam yu):
// a third of the flops of the pivoted LU, no pivot search, and half the shared memory, so a
// whole target fits one warp (no CTA barriers) and ~12 targets are resident per SM.
// Any pivot <= tol (not positive definite: product / product_sum models, unconstrained
// Nelder-Mead parameters, duplicate observations) sends the target to the pivoted-LU kernel.
//
// Storage: lower triangle as 8x8 tiles (row-major, 16-byte chunks XOR-swizzled so that the
// row-wise LDS.128 of the panel step and the mma fragment loads are both conflict-free).  Rows
// m and m+1 of the tiled matrix hold c0' and 1', so the forward substitutions ride along with
// the factorisation.
//
// Per tile column Jc (left-looking):
//   1. tiles (I, Jc), I >= Jc:  C -= sum over all earlier 4-column panels q of L(I,q) L(Jc,q)'
//      -- accumulators in registers, 2 LDS + 1 DMMA (m8n8k4) per panel and tile;
//   2. panel 2Jc (columns 8Jc..8Jc+3): every lane factors the 4x4 diagonal block redundantly in
//      registers (rsqrt = MUFU.RSQ64H + one cubic Newton step), then lanes own rows and
//      forward-substitute their 4 entries;
//   3. columns 8Jc+4..8Jc+7 of the same tiles -= panel 2Jc (one DMMA per tile, masked store);
//   4. panel 2Jc+1 likewise.

struct CholParams {
    const double4 *obs4;     // observations in ORIGINAL order: {x, y, t, resid}
    double csill;            // c
    double cdiag;            // c - gamma(0, 0)
    double tol;              // pivot tolerance
    // all-spherical fast path: C = cb + ncs q(h is) + nct q(t it) + ncm q(d im)
    double is, it, im, ncs, nct, ncm, cb, ani2;
    // the same without clamps (every neighbour pair inside every range):
    //   C = cb + h (as - bs h^2) + t (at - bt t^2) + d (am - bm d^2),  a = -c0 1.5 / r,  b = -c0 0.5 / r^3
    double pas, pbs, pat, pbt, pam, pbm;
    int clamp;               // 1: a neighbour pair can lie beyond a marginal's range
    int mode;                // assembly: 0 polynomial, 1 regrouped + clamps (or generic), 2 reference expression + fast sqrt
    const unsigned *pairtab; // pair e = a(a-1)/2 + b (b < a)  ->  a | b << 8 | chol_off(a, b) << 16
    unsigned *fail_list;     // targets (chunk-local ids) that need the pivoted LU
    unsigned *fail_count;
};

enum { CHOL_SUM_SPH = 0, CHOL_METRIC_SPH = 1, CHOL_SUM_METRIC_SPH = 2, CHOL_GENERIC = 3 };

__device__ __forceinline__ double rsqrt_approx(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// 1/sqrt(x), x > 0 and normal: MUFU.RSQ64H (2^-22) + one cubic step -> ~2^-60 relative.
__device__ __forceinline__ double rsqrt_fast(double x)
{
    const double y = rsqrt_approx(x);
    const double t = x * y;
    const double e = __fma_rn(-t, y, 1.0);
    const double p = __fma_rn(0.375, e, 0.5);     // p and ye are independent: chain = 4 ops
    const double ye = y * e;
    return __fma_rn(ye, p, y);
}

// sqrt(x) to <= 1 ulp for x >= 0 (coupled Newton on g ~ sqrt x, h ~ 1/(2 sqrt x)); the tiny
// offset keeps x = 0 (coincident stations) away from 0 * inf: sqrt(0) -> 1e-145.
__device__ __forceinline__ double sqrt_fast(double x)
{
    x = x + 1e-290;
    const double y = rsqrt_approx(x);
    double g = x * y;
    double h = 0.5 * y;
    const double r = __fma_rn(-h, g, 0.5);
    g = __fma_rn(g, r, g);
    h = __fma_rn(h, r, h);
    const double d = __fma_rn(-g, g, x);
    return __fma_rn(d, h, g);
}

// sqrt(x) for x > 0 (caller adds the guard): MUFU.RSQ64H + one cubic correction of
// t = x y ~ sqrt(x): t (1 + e/2 + 3 e^2/8), e = 1 - t y  ->  relative error ~ e^3 (2^-66) + rounding.
__device__ __forceinline__ double sqrt_cubic(double x)
{
    const double y = rsqrt_approx(x);
    const double t = x * y;
    const double e = __fma_rn(-t, y, 1.0);
    const double p = __fma_rn(0.375, e, 0.5);
    const double te = t * e;
    return __fma_rn(te, p, t);
}

// Covariance entry without clamps, polynomial form (see CholParams).
template <int VAR>
__device__ __forceinline__ double cov_entry_poly(const CholParams &X, double d2, double lt)
{
    d2 = d2 + 1e-290;                 // coincident stations: sqrt(0) -> 1e-145, no 0 * inf
    const double lt2 = lt * lt;
    double c = X.cb;
    if (VAR == CHOL_SUM_SPH || VAR == CHOL_SUM_METRIC_SPH) {
        const double h = sqrt_cubic(d2);
        c = __fma_rn(h, __fma_rn(-X.pbs, d2, X.pas), c);
        c = __fma_rn(lt, __fma_rn(-X.pbt, lt2, X.pat), c);
    }
    if (VAR == CHOL_METRIC_SPH || VAR == CHOL_SUM_METRIC_SPH) {
        const double dm2 = __fma_rn(X.ani2, lt2, d2);
        const double dm = sqrt_cubic(dm2);
        c = __fma_rn(dm, __fma_rn(-X.pbm, dm2, X.pam), c);
    }
    return c;
}

// spherical shape q(u) = 1.5 u - 0.5 u^3 for u <= 1, 1 beyond (variogram_models.py:40-44, r > 0).
// CLAMP = false when the search radii guarantee u <= 1 for every neighbour pair.
template <bool CLAMP>
__device__ __forceinline__ double sph_q(double u)
{
    if (CLAMP) u = (u < 1.0) ? u : 1.0;
    return u * __fma_rn(-0.5, u * u, 1.5);
}

template <int VAR, bool CLAMP>
__device__ __forceinline__ double cov_entry(const Model &M, const CholParams &X, double d2, double lt)
{
    if (VAR == CHOL_SUM_SPH) {
        const double h = sqrt_fast(d2);
        return __fma_rn(X.nct, sph_q<CLAMP>(lt * X.it), __fma_rn(X.ncs, sph_q<CLAMP>(h * X.is), X.cb));
    } else if (VAR == CHOL_METRIC_SPH) {
        const double dm = sqrt_fast(__fma_rn(X.ani2 * lt, lt, d2));
        return __fma_rn(X.ncm, sph_q<CLAMP>(dm * X.im), X.cb);
    } else if (VAR == CHOL_SUM_METRIC_SPH) {
        const double h = sqrt_fast(d2);
        const double dm = sqrt_fast(__fma_rn(X.ani2 * lt, lt, d2));
        double c = __fma_rn(X.ncs, sph_q<CLAMP>(h * X.is), X.cb);
        c = __fma_rn(X.nct, sph_q<CLAMP>(lt * X.it), c);
        return __fma_rn(X.ncm, sph_q<CLAMP>(dm * X.im), c);
    } else {
        return X.csill - gamma_st(M, sqrt(d2), lt);
    }
}

__host__ __device__ constexpr int chol_swz(int r) { return (((r >> 1) & 1) << 1) | ((r >> 2) & 1); }
__device__ __forceinline__ int chol_tri(int I) { return (I * (I + 1)) >> 1; }
// offset (in doubles) of entry (r, c), c <= r (or any entry of a stored tile)
__device__ __forceinline__ int chol_off(int r, int c)
{
    const int ri = r & 7, ci = c & 7;
    return (chol_tri(r >> 3) + (c >> 3)) * 64 + ri * 8 + ((((ci >> 1) ^ chol_swz(ri)) << 1) | (ci & 1));
}

// per-lane constants (computed once per kernel)
struct CholLane {
    int lane;
    int q8, l8, ls;     // row ownership: rows lane + 32 slot; q8 = lane >> 3, l8 = (lane & 7) * 8, ls = swz(lane & 7)
    int foA, foB;       // mma fragment element (row g, column t) / (row g, column 4 + t) of a tile
    int co;             // C fragment (row g, columns 2t, 2t+1)
    int t;
};

// ---- C = c - gamma(D_s, D_t), strictly lower triangle; pair e = a(a-1)/2 + b, b < a ----
// Four pairs per lane and iteration (independent dependency chains); (a, b) and the tile offset
// come from a precomputed table (coalesced, prefetched one iteration ahead) instead of a
// divergent decode loop.
template <int VAR, int MODE>
__device__ __forceinline__ double chol_cov(const Model &M, const CholParams &X, double2 pa, double ta,
                                           double2 pb, double tb)
{
    const double dx = pa.x - pb.x, dy = pa.y - pb.y;
    const double lt = fabs(ta - tb);
    if (VAR != CHOL_GENERIC && MODE == 2) {
        // the reference's expression marginal by marginal (b + c0 (...), summed in its order):
        // with large cancelling nuggets every rounding of the reference is reproduced; only the
        // square roots are the fast ones (<= 1 ulp, far below the ulp of the nuggets)
        const double h = sqrt_cubic_lu((dx * dx + dy * dy) + 1e-290);
        constexpr int ST = VAR == CHOL_SUM_SPH ? STIF_ST_SUM
                         : VAR == CHOL_METRIC_SPH ? STIF_ST_METRIC : STIF_ST_SUM_METRIC;
        return X.csill - gamma_tpl<ST, STIF_M_SPHERICAL, STIF_M_SPHERICAL, STIF_M_SPHERICAL, true>(M, h, lt);
    }
    const double d2 = __fma_rn(dy, dy, dx * dx);
    if (VAR != CHOL_GENERIC && MODE == 0) return cov_entry_poly<VAR>(X, d2, lt);
    return cov_entry<VAR, true>(M, X, d2, lt);
}

template <int VAR, int MODE>
__device__ __forceinline__ void chol_assemble(double *T, const double2 *pxy, const double *pt, int m,
                                              const Model &M, const CholParams &X, int lane)
{
    constexpr int U = 4;                       // pairs per lane and iteration
    const int npairs = (m * (m - 1)) >> 1;
    const unsigned *tab = X.pairtab + lane;
    unsigned w[U], wn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) w[u] = (lane + 32 * u < npairs) ? __ldg(tab + 32 * u) : 0x00000100u;
#pragma unroll 1
    for (int e = lane; e < npairs; e += 32 * U) {
        tab += 32 * U;
#pragma unroll
        for (int u = 0; u < U; ++u)            // next iteration's table words (pair (1,0) = harmless filler)
            wn[u] = (e + 32 * (U + u) < npairs) ? __ldg(tab + 32 * u) : 0x00000100u;
        double2 pa[U], pb[U];
        double ta[U], tb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int a = w[u] & 0xff, b = (w[u] >> 8) & 0xff;
            pa[u] = pxy[a];
            pb[u] = pxy[b];
            ta[u] = pt[a];
            tb[u] = pt[b];
        }
        double c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = chol_cov<VAR, MODE>(M, X, pa[u], ta[u], pb[u], tb[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (e + 32 * u < npairs) T[w[u] >> 16] = c[u];
            w[u] = wn[u];
        }
    }
}

// Factor the 4-column panel at columns 8 Pt + 4 H .. +3 and forward-substitute the rows below.
// Every lane factors the 4x4 diagonal block redundantly in registers; lane 0 writes it back
// (1/L_ii on the diagonal); lanes own rows lane + 32 slot.  Returns true if a pivot failed.
template <int H, int NSLOT>
__device__ __forceinline__ bool chol_panel(double *T, int Pt, int m, int nrow, double tol, const CholLane &L)
{
    const int c0 = 8 * Pt + 4 * H;
    double *Td = T + (chol_tri(Pt) + Pt) * 64;
    constexpr int r0 = 4 * H;
    const int ncol = m - c0;   // real columns in this panel (>= 1); the others pass through
    const int rstart = c0 + (ncol < 4 ? ncol : 4);
    // own rows (lane + 32 s) below the real part of the block: loaded up front so that their
    // latency hides behind the serial 4x4 factorisation
    const int chA = ((2 * H) ^ L.ls) << 1, chB = ((2 * H + 1) ^ L.ls) << 1;
    double2 q0[NSLOT], q1[NSLOT];
    double2 *p0[NSLOT];
    bool act[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int r = L.lane + 32 * s;
        act[s] = r >= rstart && r < nrow;
        double *row = T + (chol_tri(L.q8 + 4 * s) + Pt) * 64 + L.l8;
        p0[s] = reinterpret_cast<double2 *>(row + chA);
        if (32 * s + 31 >= rstart && 32 * s < nrow) {      // warp-uniform: slot has rows to do
            if (act[s]) {
                q0[s] = *p0[s];
                q1[s] = *reinterpret_cast<double2 *>(row + chB);
            }
        }
    }
    // 16-byte chunk ch of row ri of the tile (swizzle is a compile-time constant here)
#define CHUNK(ri, ch) reinterpret_cast<double2 *>(Td + (ri) * 8 + ((((ch)) ^ chol_swz(ri)) << 1))
    const double2 u0 = *CHUNK(r0 + 0, 2 * H);
    const double2 u1 = *CHUNK(r0 + 1, 2 * H);
    const double2 u2 = *CHUNK(r0 + 2, 2 * H), v2 = *CHUNK(r0 + 2, 2 * H + 1);
    const double2 u3 = *CHUNK(r0 + 3, 2 * H), v3 = *CHUNK(r0 + 3, 2 * H + 1);
    double d00 = u0.x, d10 = u1.x, d11 = u1.y, d20 = u2.x, d21 = u2.y, d22 = v2.x;
    double d30 = u3.x, d31 = u3.y, d32 = v3.x, d33 = v3.y;
    __syncwarp();              // every lane holds the block before it is overwritten
    if (ncol < 4) {
        d30 = 0.0; d31 = 0.0; d32 = 0.0; d33 = 1.0;
        if (ncol < 3) { d20 = 0.0; d21 = 0.0; d22 = 1.0; }
        if (ncol < 2) { d10 = 0.0; d11 = 1.0; }
    }
    bool fail = !(d00 > tol);
    const double i0 = rsqrt_fast(d00);
    const double l10 = d10 * i0, l20 = d20 * i0, l30 = d30 * i0;
    d11 = __fma_rn(-l10, l10, d11);
    d21 = __fma_rn(-l20, l10, d21);
    d31 = __fma_rn(-l30, l10, d31);
    d22 = __fma_rn(-l20, l20, d22);
    d32 = __fma_rn(-l30, l20, d32);
    d33 = __fma_rn(-l30, l30, d33);
    fail |= !(d11 > tol);
    const double i1 = rsqrt_fast(d11);
    const double l21 = d21 * i1, l31 = d31 * i1;
    d22 = __fma_rn(-l21, l21, d22);
    d32 = __fma_rn(-l31, l21, d32);
    d33 = __fma_rn(-l31, l31, d33);
    fail |= !(d22 > tol);
    const double i2 = rsqrt_fast(d22);
    const double l32 = d32 * i2;
    d33 = __fma_rn(-l32, l32, d33);
    fail |= !(d33 > tol);
    const double i3 = rsqrt_fast(d33);
    if (fail) return true;
    if (L.lane == 0) {         // the real rows of the diagonal block
        CHUNK(r0 + 0, 2 * H)->x = i0;
        if (ncol >= 2) *CHUNK(r0 + 1, 2 * H) = make_double2(l10, i1);
        if (ncol >= 3) {
            *CHUNK(r0 + 2, 2 * H) = make_double2(l20, l21);
            CHUNK(r0 + 2, 2 * H + 1)->x = i2;
        }
        if (ncol >= 4) {
            *CHUNK(r0 + 3, 2 * H) = make_double2(l30, l31);
            *CHUNK(r0 + 3, 2 * H + 1) = make_double2(l32, i3);
        }
    }
#undef CHUNK
    // forward substitution of the preloaded rows (pass-through columns have i = 1, l = 0)
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        if (32 * s + 31 >= rstart && 32 * s < nrow) {
            if (act[s]) {
                const double x0 = q0[s].x * i0;
                const double x1 = __fma_rn(-x0, l10, q0[s].y) * i1;
                const double x2 = __fma_rn(-x1, l21, __fma_rn(-x0, l20, q1[s].x)) * i2;
                const double x3 = __fma_rn(-x2, l32, __fma_rn(-x1, l31, __fma_rn(-x0, l30, q1[s].y))) * i3;
                *p0[s] = make_double2(x0, x1);
                *reinterpret_cast<double2 *>(reinterpret_cast<double *>(p0[s]) - chA + chB) = make_double2(x2, x3);
            }
        }
    }
    return false;
}

// tiles (I0 .. I0+NTI-1, Jc) -= sum over the 2 Jc earlier panels of L(I, q) L(Jc, q)^T
template <int NTI>
__device__ __forceinline__ void chol_update_group(double *T, int I0, int Jc, const CholLane &L)
{
    double2 acc[NTI];
    const double *pa[NTI];
#pragma unroll
    for (int u = 0; u < NTI; ++u) {
        acc[u] = make_double2(0.0, 0.0);
        pa[u] = T + chol_tri(I0 + u) * 64;
    }
    const double *pb = T + chol_tri(Jc) * 64;
#pragma unroll 1
    for (int qq = 0; qq < Jc; ++qq) {
        const double bA = pb[L.foA], bB = pb[L.foB];
#pragma unroll
        for (int u = 0; u < NTI; ++u) {
            const double aA = pa[u][L.foA], aB = pa[u][L.foB];
            dmma884(acc[u].x, acc[u].y, aA, bA);
            dmma884(acc[u].x, acc[u].y, aB, bB);
            pa[u] += 64;
        }
        pb += 64;
    }
#pragma unroll
    for (int u = 0; u < NTI; ++u) {     // pa[u] now points at tile (I0 + u, Jc)
        double2 *cp = reinterpret_cast<double2 *>(const_cast<double *>(pa[u]) + L.co);
        double2 cv = *cp;
        cv.x -= acc[u].x;
        cv.y -= acc[u].y;
        *cp = cv;
    }
}

// columns 4..7 of tiles (I0 .. I0+NTI-1, Jc) -= left panel of the same tile column
template <int NTI>
__device__ __forceinline__ void chol_half_group(double *T, int I0, int Jc, const CholLane &L)
{
    const double bA = T[(chol_tri(Jc) + Jc) * 64 + L.foA];
    double2 acc[NTI];
    double *pc[NTI];
#pragma unroll
    for (int u = 0; u < NTI; ++u) {
        pc[u] = T + (chol_tri(I0 + u) + Jc) * 64;
        acc[u] = make_double2(0.0, 0.0);
        dmma884(acc[u].x, acc[u].y, pc[u][L.foA], bA);
    }
    if (L.t >= 2) {
#pragma unroll
        for (int u = 0; u < NTI; ++u) {
            double2 *cp = reinterpret_cast<double2 *>(pc[u] + L.co);
            double2 cv = *cp;
            cv.x -= acc[u].x;
            cv.y -= acc[u].y;
            *cp = cv;
        }
    }
}

constexpr int CHOL_G = 8;   // tiles of one tile column updated together (accumulators in registers)

#define CHOL_DISPATCH_NTI(FN, n_, ...)                                                             \
    switch (n_) {                                                                                  \
    case 1: FN<1>(__VA_ARGS__); break;                                                             \
    case 2: FN<2>(__VA_ARGS__); break;                                                             \
    case 3: FN<3>(__VA_ARGS__); break;                                                             \
    case 4: FN<4>(__VA_ARGS__); break;                                                             \
    case 5: FN<5>(__VA_ARGS__); break;                                                             \
    case 6: FN<6>(__VA_ARGS__); break;                                                             \
    case 7: FN<7>(__VA_ARGS__); break;                                                             \
    default: FN<8>(__VA_ARGS__); break;                                                            \
    }

// One tile row (8 steps) of the back substitution L' w = z; NS = column slots that can still be
// below the diagonal.  z[i] belongs to column lane + 32 i; w_j goes to sol[j].
template <int ZR, int NS>
__device__ __forceinline__ void chol_backsub_row(const double *T, double *sol, double (&z)[ZR], int J,
                                                 int m, const CholLane &L)
{
    const double *rowT = T + chol_tri(J) * 64;
    // lane's column offsets for the 4 possible row swizzles
    int lo[NS][4];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const int c = L.lane + 32 * i;
#pragma unroll
        for (int s = 0; s < 4; ++s) lo[i][s] = (c >> 3) * 64 + (((((c & 7) >> 1) ^ s) << 1) | (c & 1));
    }
    const int zi = (8 * J) >> 5;          // slot that owns this tile row's columns
    const int src0 = (8 * J) & 31;
    double wkeep = 0.0;                   // lane jj keeps w_{8J+jj}; stored once after the 8 steps
#pragma unroll
    for (int jj = 7; jj >= 0; --jj) {
        const int j = 8 * J + jj;
        if (j >= m) continue;             // padding / right-hand-side rows of the last tile row
        const int sj = chol_swz(jj);
        const double dinv = rowT[J * 64 + jj * 8 + ((((jj >> 1) ^ sj) << 1) | (jj & 1))];
        double zj = z[0];
#pragma unroll
        for (int i = 1; i < NS; ++i)
            if (zi == i) zj = z[i];
        zj = __shfl_sync(FULL, zj, src0 + jj);
        const double wj = zj * dinv;
        if (L.lane == jj) wkeep = wj;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int c = L.lane + 32 * i;
            if (c < j) z[i] = __fma_rn(-rowT[jj * 8 + lo[i][sj]], wj, z[i]);
        }
    }
    if (L.lane < 8 && 8 * J + L.lane < m) sol[8 * J + L.lane] = wkeep;
}

template <int VAR, int ZR, int MODE>
__global__ void __maxnreg__(ZR == 2 ? 152 : 240) solve_chol_kernel(SolveParams P, CholParams X)
{
    extern __shared__ __align__(16) double smem_d[];
    const int lane = threadIdx.x;
    const int k = P.k;
    const int kp = (k + 1) & ~1;
    const int ntmax = (k + 2 + 7) >> 3;
    double *T = smem_d;
    double2 *pxy = reinterpret_cast<double2 *>(T + chol_tri(ntmax) * 64);
    double *pt = reinterpret_cast<double *>(pxy + kp);
    double *g0 = pt + kp, *rs = g0 + kp;
    unsigned *ids = reinterpret_cast<unsigned *>(rs + kp);
    double *sol = pt;                                   // pt is dead after the assembly
    CholLane L;
    {
        const int g = lane >> 2, t = lane & 3, sg = chol_swz(g);
        L.lane = lane;
        L.q8 = lane >> 3;
        L.l8 = (lane & 7) * 8;
        L.ls = chol_swz(lane & 7);
        L.foA = g * 8 + ((((t >> 1) ^ sg) << 1) | (t & 1));
        L.foB = g * 8 + ((((2 + (t >> 1)) ^ sg) << 1) | (t & 1));
        L.co = g * 8 + ((t ^ sg) << 1);
        L.t = t;
    }

    // The next target's selection (indices, then coordinates) is fetched into registers while
    // the current one is being factorised: no dependent global-load chain at the loop top.
    unsigned tl = 0, ntl = 0, nid[ZR];
    int m = 0, nm = 0;
    double4 no[ZR];
    double ng[ZR];
#pragma unroll
    for (int i = 0; i < ZR; ++i) {
        nid[i] = 0u;
        no[i] = make_double4(0.0, 0.0, 0.0, 0.0);
        ng[i] = 0.0;
    }
    bool has_next = false;
    auto fetch_ids = [&]() {              // stage B: needs ntl
        nm = has_next ? P.sel_n[ntl] : 0;
#pragma unroll
        for (int i = 0; i < ZR; ++i) {
            const int e = lane + 32 * i;
            if (has_next && e < k) nid[i] = P.sel_idx[(size_t)ntl * k + e];
        }
    };
    auto fetch_obs = [&]() {              // stage C: needs nm and nid
#pragma unroll
        for (int i = 0; i < ZR; ++i) {
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
        __syncwarp();
        // ---- neighbours (prefetched), diagonal, the two right-hand-side rows ----
        if (m > 0) {
#pragma unroll
            for (int i = 0; i < ZR; ++i) {
                const int e = lane + 32 * i;
                if (e < m) {
                    ids[e] = nid[i];
                    pxy[e] = make_double2(no[i].x, no[i].y);
                    pt[e] = no[i].z;
                    rs[e] = no[i].w;
                    g0[e] = ng[i];
                    T[chol_off(e, e)] = X.cdiag;
                    T[chol_off(m, e)] = X.csill - ng[i];
                    T[chol_off(m + 1, e)] = 1.0;
                }
            }
        }
        has_next = slot + (int)gridDim.x < P.n;
        if (has_next) ntl = P.order[slot + gridDim.x];       // stage A
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
        const int nrow = m + 2, nt = (nrow + 7) >> 3;
        __syncwarp();
        if (m > 1) chol_assemble<VAR, MODE>(T, pxy, pt, m, P.M, X, lane);
        __syncwarp();
        fetch_ids();
        // ---- left-looking tiled Cholesky ----
        bool fail = false;
        const int ntc = (m + 7) >> 3;       // tile columns that hold real columns
        for (int Jc = 0; Jc < ntc; ++Jc) {
            if (Jc > 0) {
                for (int I0 = Jc; I0 < nt; I0 += CHOL_G) {
                    CHOL_DISPATCH_NTI(chol_update_group, nt - I0, T, I0, Jc, L);
                }
                __syncwarp();
            }
            fail = chol_panel<0, ZR + 1>(T, Jc, m, nrow, X.tol, L);
            __syncwarp();
            if (fail || 8 * Jc + 4 >= m) break;
            for (int I0 = Jc; I0 < nt; I0 += CHOL_G) {
                CHOL_DISPATCH_NTI(chol_half_group, nt - I0, T, I0, Jc, L);
            }
            __syncwarp();
            fail = chol_panel<1, ZR + 1>(T, Jc, m, nrow, X.tol, L);
            __syncwarp();
            if (fail) break;
        }
        fetch_obs();
        if (fail) {   // not positive definite / singular: the pivoted LU takes this target
            if (lane == 0) X.fail_list[atomicAdd(X.fail_count, 1u)] = tl;
            continue;
        }
        // ---- lam and z = yv + lam yu (rows m, m+1 now hold yv', yu') ----
        double z[ZR];
        {
            double yu[ZR], suu = 0.0, suv = 0.0;
#pragma unroll
            for (int i = 0; i < ZR; ++i) {
                const int c = lane + 32 * i;
                yu[i] = 0.0;
                z[i] = 0.0;
                if (c < m) {
                    z[i] = T[chol_off(m, c)];
                    yu[i] = T[chol_off(m + 1, c)];
                    suu = __fma_rn(yu[i], yu[i], suu);
                    suv = __fma_rn(yu[i], z[i], suv);
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                suu += __shfl_xor_sync(FULL, suu, off);
                suv += __shfl_xor_sync(FULL, suv, off);
            }
            const double lam = (1.0 - suv) / suu;
#pragma unroll
            for (int i = 0; i < ZR; ++i) z[i] = __fma_rn(lam, yu[i], z[i]);
        }
        __syncwarp();          // pt (aliased by sol) is dead, rows m / m+1 have been read
        // ---- back substitution L' w = z by tile rows (padding columns of the last one: z = 0,
        //      their dinv entries are finite garbage-free pass-through values or unused) ----
        for (int J = ntc - 1; J >= 0; --J) {
            const int ns = ((8 * J + 7) >> 5) + 1;
            if (ZR == 2) {
                if (ns == 1) chol_backsub_row<ZR, 1>(T, sol, z, J, m, L);
                else chol_backsub_row<ZR, 2>(T, sol, z, J, m, L);
            } else {
                if (ns == 1) chol_backsub_row<ZR, 1>(T, sol, z, J, m, L);
                else if (ns == 2) chol_backsub_row<ZR, 2>(T, sol, z, J, m, L);
                else if (ns == 3) chol_backsub_row<ZR, 3>(T, sol, z, J, m, L);
                else chol_backsub_row<ZR, 4>(T, sol, z, J, m, L);
            }
        }
        __syncwarp();
        solve_epilogue(P, sol, g0, ids, rs, reinterpret_cast<double *>(pxy),
                       reinterpret_cast<float *>(pxy + (kp >> 1)),
                       reinterpret_cast<float *>(pxy + (kp >> 1)) + kp, m, tg, lane);
    }
}
