// kriging.cu -- space-time ordinary kriging on sm_100a.
//
// Replaces (reference file:line):
//   stif/data.py:170-212        Data.get_kriging_idxs  (dense n_obs x batch pre-filter)
//   stif/predictor.py:602-660   nd_kriging             (candidate gamma-vector, top-k)
//   stif/predictor.py:566-593   calc_kriging_weights   (bordered system + lstsq)
//   stif/predictor.py:547-556   variogram_model        (+ stif/variogram_models.py:7-218)
//   stif/predictor.py:882-906   mean / std epilogue
//
// Pipeline per call (one stream, no host synchronisation inside):
//   grid   : observations binned into a uniform space-time grid with two cells per search
//            radius; (cell, obs-index) radix-sorted (CUB, plumbing), so each cell is a contiguous
//            run in ascending observation index.  Cached per (obs set, radii).
//   search : one warp per target (targets visited in grid-cell order for L2 locality).  The
//            causal 5x5x3 cell stencil is 15 contiguous runs, concatenated into one virtual
//            sequence; lanes stream it coalesced (four loads in flight), apply the reference's
//            exact filter
//              rn(dx*dx)+rn(dy*dy) < smax^2  &  |dt| < tmax  &  t_obs <= t_tgt  &  not left out
//            (stif/data.py:195-211; unfused like scipy cdist), compact survivors with a
//            ballot, evaluate gamma(h,t) densely, and keep the k lowest (gamma, index) with a
//            warp-wide select; bounded shared memory for any candidate count.
//   solve  : valid variogram models: covariance-form Cholesky, one warp per target
//            (krig_chol.cuh); every other model, and targets whose covariance matrix is not
//            positive definite: pivoted LU of the reference's bordered system, one CTA per
//            target (solve_blk_kernel).  Both end in the fused epilogue (float32 storage
//            emulation + numpy pairwise order, or all-float64).
// Also here: the stand-alone neighbour pair list (Data.get_kriging_idxs), the elementwise
// model evaluation, and the objective of the variogram-model fit.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <math.h>

namespace stif {
namespace krig {

constexpr int KMAX = 128;        // max_kriging_points limit
#ifndef LU_WPT_MID
#define LU_WPT_MID 2              // warps per target of the pivoted LU, 32 <= k <= 64
#endif
#ifndef LU_WPT_BIG
#define LU_WPT_BIG 4              // ... k > 64
#endif
constexpr int SEARCH_WARPS = 4;  // warps per block in the search kernel
constexpr int SCAN_U = 4;        // independent 32-candidate loads in flight per warp
constexpr int PEND_CAP = 160;    // pending in-window candidates per warp; >= 31 + 32 * SCAN_U
constexpr int BUF_CAP = 320;     // (gamma, idx) buffer per warp; >= KMAX + PEND_CAP
constexpr unsigned FULL = 0xffffffffu;

struct Model {
    int st;                 // STIF_ST_*
    int ks, kt, km;         // marginal kinds
    double rs, c0s, bs, is; // space marginal: range, sill, nugget, inverse (1/r or 1/(r/2)^2)
    double rt, c0t, bt, it; // time marginal
    double rm, c0m, bm, im; // metric marginal
    double ani2;            // ani*ani
    double k;               // product_sum weighting coefficient
};

struct Grid {
    const double4 *obs;     // cell-sorted {x, y, t, bits(idx)}
    const unsigned *cell_start;  // [ncells + 1]
    double x0, y0, t0;
    double inv_cx, inv_cy, inv_ct;
    int nx, ny, nt;
};

// ---- variogram model (stif/variogram_models.py) -----------------------------------
// Formulas are evaluated verbatim (no validity assumption: Nelder-Mead parameters are
// unconstrained); h/a is evaluated as h*(1/a) like the reference's fast-math code.
__device__ __forceinline__ double marginal(int kind, double h, double r, double c0, double b,
                                           double inv)
{
    if (kind == STIF_M_SPHERICAL) {  // variogram_models.py:40-44
        if (h <= r) {
            const double u = h * inv;
            return b + c0 * ((1.5 * u) - (0.5 * ((u * u) * u)));
        }
        return b + c0;
    }
    // gaussian, variogram_models.py:70-71: a = r/2; b + c0*(1 - exp(-(h^2/a^2)))
    return b + c0 * (1.0 - exp(-((h * h) * inv)));
}

// sqrt for the assembly fast paths: MUFU.RSQ64H + one cubic correction (<= ~1 ulp), x > 0
__device__ __forceinline__ double sqrt_cubic_lu(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = __fma_rn(-t, y, 1.0);
    const double p = __fma_rn(0.375, e, 0.5);
    return __fma_rn(t * e, p, t);
}

// ST / KS / KT / KM < 0: read the id from the model at run time; >= 0: compile-time constant
// (same arithmetic, the warp-uniform branches disappear).  FSQ: the metric distance uses the
// fast square root (<= 1 ulp) instead of the IEEE one -- for the kriging MATRIX only, never
// for the gamma-vector that decides the neighbour set.
template <int ST, int KS, int KT, int KM, bool FSQ = false>
__device__ __forceinline__ double gamma_tpl(const Model &M, double h, double t)
{
    const int st = ST < 0 ? M.st : ST;
    const int ks = KS < 0 ? M.ks : KS, kt = KT < 0 ? M.kt : KT, km = KM < 0 ? M.km : KM;
    switch (st) {
    case STIF_ST_SUM:  // :98-99
        return marginal(ks, h, M.rs, M.c0s, M.bs, M.is) + marginal(kt, t, M.rt, M.c0t, M.bt, M.it);
    case STIF_ST_PRODUCT:  // :126-127
        return marginal(ks, h, M.rs, M.c0s, M.bs, M.is) * marginal(kt, t, M.rt, M.c0t, M.bt, M.it);
    case STIF_ST_PRODUCT_SUM: {  // :155-157  k*prod + prod
        const double p = marginal(ks, h, M.rs, M.c0s, M.bs, M.is) *
                         marginal(kt, t, M.rt, M.c0t, M.bt, M.it);
        return M.k * p + p;
    }
    case STIF_ST_METRIC: {  // :184-186
        const double d2m = h * h + (M.ani2 * t) * t;
        const double d = FSQ ? sqrt_cubic_lu(d2m + 1e-290) : sqrt(d2m);
        return marginal(km, d, M.rm, M.c0m, M.bm, M.im);
    }
    default: {  // sum_metric :213-218
        const double d2m = h * h + (M.ani2 * t) * t;
        const double d = FSQ ? sqrt_cubic_lu(d2m + 1e-290) : sqrt(d2m);
        return (marginal(ks, h, M.rs, M.c0s, M.bs, M.is) +
                marginal(kt, t, M.rt, M.c0t, M.bt, M.it)) +
               marginal(km, d, M.rm, M.c0m, M.bm, M.im);
    }
    }
}

__device__ __forceinline__ double gamma_st(const Model &M, double h, double t)
{
    return gamma_tpl<-1, -1, -1, -1>(M, h, t);
}

// search-kernel variants: 0 = any model (run-time ids), 1 = sum_metric, 2 = product_sum, both
// with three spherical marginals (the reference's defaults, predictor.py:662-666)
template <int GV>
__device__ __forceinline__ double gamma_search(const Model &M, double h, double t)
{
    if (GV == 1) return gamma_tpl<STIF_ST_SUM_METRIC, STIF_M_SPHERICAL, STIF_M_SPHERICAL, STIF_M_SPHERICAL>(M, h, t);
    if (GV == 2) return gamma_tpl<STIF_ST_PRODUCT_SUM, STIF_M_SPHERICAL, STIF_M_SPHERICAL, STIF_M_SPHERICAL>(M, h, t);
    return gamma_st(M, h, t);
}

__global__ void model_eval_kernel(Model M, const double *h, const double *t, int64_t n, double *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = gamma_st(M, h[i], t[i]);
}

// ---- grid build --------------------------------------------------------------------
__global__ void bbox_kernel(const double *x, const double *y, const double *t, int64_t n,
                            double *bbox /* [6]: min x,y,t, max x,y,t; pre-initialised */)
{
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double v[3] = {x[i], y[i], t[i]};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (v[d] < mn[d]) mn[d] = v[d];   // NaN never updates
            if (v[d] > mx[d]) mx[d] = v[d];
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        for (int off = 16; off > 0; off >>= 1) {
            mn[d] = fmin(mn[d], __shfl_xor_sync(FULL, mn[d], off));
            mx[d] = fmax(mx[d], __shfl_xor_sync(FULL, mx[d], off));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        // atomic min/max on doubles via ordered integer keys
        auto enc = [](double v) -> unsigned long long {
            unsigned long long b = (unsigned long long)__double_as_longlong(v);
            return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
        };
        unsigned long long *bb = reinterpret_cast<unsigned long long *>(bbox);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&bb[d], enc(mn[d]));
            atomicMax(&bb[3 + d], enc(mx[d]));
        }
    }
}

__device__ __forceinline__ double dec_key(unsigned long long k)
{
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

struct GridDims {
    double x0, y0, t0, inv_cx, inv_cy, inv_ct;
    int nx, ny, nt;
    int hx, hy, ht;   // stencil half-widths in cells: ceil(radius / cell)
};

constexpr double GRID_SUB = 2.0;   // cells per search radius (finer cells: fewer candidates examined)
constexpr double GRID_SUB_X = 8.0; // ... along x, where the cells of a stencil row are one contiguous run

// Decide grid geometry on the device (no host round trip): cells are >= radius*(1+1e-9)
// wide and the total number of cells is capped at max_cells by widening them.
__global__ void grid_dims_kernel(const double *bbox_enc, double smax, double tmax, int64_t max_cells,
                                 double sub_x, double sub_y, double sub_t, GridDims *out)
{
    const unsigned long long *bb = reinterpret_cast<const unsigned long long *>(bbox_enc);
    double lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {
        lo[d] = dec_key(bb[d]);
        hi[d] = dec_key(bb[3 + d]);
        if (!(hi[d] >= lo[d])) {  // empty or all-NaN
            lo[d] = 0.0;
            hi[d] = 0.0;
        }
    }
    const double radius[3] = {smax, smax, tmax};
    const double sub[3] = {sub_x, sub_y, sub_t};
    double cell[3], n[3];
    bool valid[3];
    for (int d = 0; d < 3; ++d) {
        valid[d] = radius[d] > 0.0 && !isinf(radius[d]);
        cell[d] = valid[d] ? radius[d] / sub[d] : (hi[d] - lo[d]) + 1.0;
        cell[d] *= (1.0 + 1e-9);
        n[d] = floor((hi[d] - lo[d]) / cell[d]) + 2.0;  // one spare cell at the upper edge
        if (!(n[d] >= 2.0)) n[d] = 2.0;
    }
    // the x axis may be finer than the others (its cells are contiguous in memory, so a finer x
    // costs no extra runs): halve its subdivision while the grid does not fit the cap
    for (double sx = sub_x; valid[0] && sx > sub_y && n[0] * n[1] * n[2] > (double)max_cells;) {
        sx *= 0.5;
        if (sx < sub_y) sx = sub_y;
        cell[0] = radius[0] / sx * (1.0 + 1e-9);
        n[0] = floor((hi[0] - lo[0]) / cell[0]) + 2.0;
        if (!(n[0] >= 2.0)) n[0] = 2.0;
    }
    // cap: widen all cells by a common factor until the product fits
    for (int it = 0; it < 64 && n[0] * n[1] * n[2] > (double)max_cells; ++it) {
        for (int d = 0; d < 3; ++d) {
            if (n[d] > 2.0) {
                cell[d] *= 1.26;
                n[d] = floor((hi[d] - lo[d]) / cell[d]) + 2.0;
            }
        }
    }
    out->x0 = lo[0];
    out->y0 = lo[1];
    out->t0 = lo[2];
    out->inv_cx = 1.0 / cell[0];
    out->inv_cy = 1.0 / cell[1];
    out->inv_ct = 1.0 / cell[2];
    out->nx = (int)n[0];
    out->ny = (int)n[1];
    out->nt = (int)n[2];
    int h[3];
    for (int d = 0; d < 3; ++d) {
        double hd = valid[d] ? ceil(radius[d] / cell[d]) : n[d];
        if (!(hd >= 1.0)) hd = 1.0;
        if (hd > n[d]) hd = n[d];
        h[d] = (int)hd;
    }
    out->hx = h[0];
    out->hy = h[1];
    out->ht = h[2];
}

// cell index, clamped to [-h-1, n+h] so that a stencil of half-width h around a point far
// outside the grid is empty
__device__ __forceinline__ int raw_cell(double v, double v0, double inv, int n, int h = 1)
{
    double c = floor((v - v0) * inv);
    const double lo = -(double)h - 1.0, hi = (double)(n + h);
    if (!(c >= lo)) c = lo;  // also catches NaN
    if (c > hi) c = hi;
    return (int)c;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void obs_keys_kernel(const double *x, const double *y, const double *t, int64_t n,
                                const GridDims *gd, unsigned *keys, unsigned *idx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GridDims g = *gd;
    const int cx = clampi(raw_cell(x[i], g.x0, g.inv_cx, g.nx), 0, g.nx - 1);
    const int cy = clampi(raw_cell(y[i], g.y0, g.inv_cy, g.ny), 0, g.ny - 1);
    const int ct = clampi(raw_cell(t[i], g.t0, g.inv_ct, g.nt), 0, g.nt - 1);
    keys[i] = (unsigned)((ct * g.ny + cy) * g.nx + cx);
    idx[i] = (unsigned)i;
}

__global__ void obs_gather_kernel(const double *x, const double *y, const double *t,
                                  const unsigned *idx_sorted, int64_t n, double4 *out)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned i = idx_sorted[k];
    out[k] = make_double4(x[i], y[i], t[i], __longlong_as_double((long long)i));
}

// cell_start[c] = first sorted position with key >= c  (c in [0, ncells]); one thread per cell.
__global__ void cell_start_kernel(const unsigned *keys_sorted, int64_t n, const GridDims *gd,
                                  int64_t max_cells, unsigned *cell_start)
{
    const GridDims g = *gd;
    const int64_t ncells = (int64_t)g.nx * g.ny * g.nt;
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncells || c > max_cells) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)keys_sorted[mid] < c) lo = mid + 1; else hi = mid;
    }
    cell_start[c] = (unsigned)lo;
}

__global__ void target_keys_kernel(const double *tx, const double *ty, const double *tt, int64_t n,
                                   const GridDims *gd, unsigned *keys, unsigned *idx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GridDims g = *gd;
    const int cx = clampi(raw_cell(tx[i], g.x0, g.inv_cx, g.nx), 0, g.nx - 1);
    const int cy = clampi(raw_cell(ty[i], g.y0, g.inv_cy, g.ny), 0, g.ny - 1);
    const int ct = clampi(raw_cell(tt[i], g.t0, g.inv_ct, g.nt), 0, g.nt - 1);
    keys[i] = (unsigned)((ct * g.ny + cy) * g.nx + cx);
    idx[i] = (unsigned)i;
}

// ---- neighbour search + top-k ------------------------------------------------------
__device__ __forceinline__ unsigned long long gamma_key(double g)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(g);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double key_gamma(unsigned long long kb)
{
    kb = (kb & 0x8000000000000000ull) ? (kb & 0x7fffffffffffffffull) : ~kb;
    return __longlong_as_double((long long)kb);
}

struct WarpBuf {
    unsigned long long *key;  // [BUF_CAP]
    unsigned *idx;            // [BUF_CAP]
    double *pend_d2;          // [PEND_CAP] squared spatial distance of a pending candidate
    double *pend_lt;          // [PEND_CAP] |dt|
    unsigned *pend_idx;       // [PEND_CAP] observation index
};

__device__ __forceinline__ int warp_count_le(const unsigned long long *key, int m, int lane,
                                             unsigned long long thr)
{
    int c = 0;
#pragma unroll
    for (int i = 0; i < BUF_CAP / 32; ++i) {       // m <= BUF_CAP: fixed trip count, predicated loads
        const int e = lane + 32 * i;
        if (e < m) c += (key[e] <= thr) ? 1 : 0;
    }
    return __reduce_add_sync(FULL, c);
}

// Keep the k smallest (key, idx) of buf[0..m) (unsorted, compacted to the front).
__device__ void select_k(const WarpBuf &B, int &m, int k, int lane)
{
    unsigned long long lo = ~0ull, hi = 0ull;
#pragma unroll
    for (int i = 0; i < BUF_CAP / 32; ++i) {
        const int e = lane + 32 * i;
        if (e < m) {
            const unsigned long long v = B.key[e];
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long a = __shfl_xor_sync(FULL, lo, off);
        const unsigned long long b = __shfl_xor_sync(FULL, hi, off);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    // k-th smallest key: regula falsi on the gamma values (the counts are a smooth function of
    // the threshold, so the interpolated guess lands within a few candidates) alternating with
    // plain bisection of the key range (guaranteed progress); early exit when a threshold
    // splits exactly k.  Invariant: the k-th smallest key is in [lo, hi],
    // c_lo = count(key < lo) < k <= c_hi = count(key <= hi).
    unsigned long long thr = hi;
    bool exact = false;
    int c_lo = 0, c_hi = m;
    for (int it = 0; lo < hi; ++it) {
        unsigned long long mid = lo + ((hi - lo) >> 1);
        if ((it & 1) == 0) {
            const double dlo = key_gamma(lo), dhi = key_gamma(hi);
            const double frac = ((double)(k - c_lo) - 0.5) / (double)(c_hi - c_lo);
            const double g = dlo + (dhi - dlo) * frac;
            if (g >= dlo && g < dhi) {          // finite, inside the bracket
                const unsigned long long gk = gamma_key(g);
                if (gk >= lo && gk < hi) mid = gk;
            }
        }
        const int c = warp_count_le(B.key, m, lane, mid);
        if (c == k) {
            thr = mid;
            exact = true;
            break;
        }
        if (c > k) {
            hi = mid;
            c_hi = c;
        } else {
            lo = mid + 1;
            c_lo = c;
        }
    }
    unsigned ithr = 0xffffffffu;
    if (!exact) {
        thr = lo;  // k-th smallest key value; ties at thr are broken by lowest index
        int c_less = 0, c_tie = 0;
        for (int e = lane; e < m; e += 32) {
            const unsigned long long v = B.key[e];
            c_less += (v < thr) ? 1 : 0;
            c_tie += (v == thr) ? 1 : 0;
        }
        c_less = __reduce_add_sync(FULL, c_less);
        c_tie = __reduce_add_sync(FULL, c_tie);
        const int need = k - c_less;  // 1 <= need <= c_tie
        if (need < c_tie) {
            unsigned ilo = 0u, ihi = 0xffffffffu;  // smallest i with count(tie & idx <= i) >= need
            while (ilo < ihi) {
                const unsigned mid = ilo + ((ihi - ilo) >> 1);
                int c = 0;
                for (int e = lane; e < m; e += 32) c += (B.key[e] == thr && B.idx[e] <= mid) ? 1 : 0;
                c = __reduce_add_sync(FULL, c);
                if (c >= need) ihi = mid; else ilo = mid + 1;
            }
            ithr = ilo;
        }
    }
    // in-place compaction (writes never pass the chunk being read)
    int out = 0;
    for (int base = 0; base < m; base += 32) {
        const int e = base + lane;
        unsigned long long v = 0;
        unsigned id = 0;
        bool sel = false;
        if (e < m) {
            v = B.key[e];
            id = B.idx[e];
            sel = exact ? (v <= thr) : (v < thr || (v == thr && id <= ithr));
        }
        const unsigned bal = __ballot_sync(FULL, sel);
        __syncwarp();
        if (sel) {
            const int pos = out + __popc(bal & ((1u << lane) - 1u));
            B.key[pos] = v;
            B.idx[pos] = id;
        }
        out += __popc(bal);
        __syncwarp();
    }
    m = out;  // == k
}

// Sort buf[0..m) by (use_key ? key : 0, idx) ascending; m <= KMAX.
// m <= 64: bitonic network in registers (two entries per lane, partner exchange by shuffle);
// larger: bitonic network in shared memory.  Padding entries sort last.
__device__ void sort_buf(const WarpBuf &B, int m, bool use_key, int lane)
{
    int p2 = 2;
    while (p2 < m) p2 <<= 1;
    if (p2 <= 64) {
        // sort key K (64 bit) with tie-breaker / payload I: by gamma -> (key, idx); by index ->
        // (idx, position) and the gamma keys are gathered afterwards
        unsigned long long K[2];
        unsigned I[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int e = 2 * lane + r;
            if (e < m) {
                const unsigned id = B.idx[e];
                K[r] = use_key ? B.key[e] : (unsigned long long)id;
                I[r] = use_key ? id : (unsigned)e;
            } else {
                K[r] = ~0ull;
                I[r] = 0xffffffffu;
            }
        }
        for (int size = 2; size <= p2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                if (stride == 1) {            // partner is the lane's other entry
                    const bool up = (((2 * lane) & size) == 0);
                    const bool gt = (K[0] > K[1]) || (K[0] == K[1] && I[0] > I[1]);
                    const bool sw = (gt == up);
                    const unsigned long long k0 = sw ? K[1] : K[0], k1 = sw ? K[0] : K[1];
                    const unsigned i0 = sw ? I[1] : I[0], i1 = sw ? I[0] : I[1];
                    K[0] = k0;
                    K[1] = k1;
                    I[0] = i0;
                    I[1] = i1;
                } else {
                    const int d = stride >> 1;             // partner lane distance
                    const bool lower = (lane & d) == 0;
                    const bool up = (((2 * lane) & size) == 0);
                    const bool keep_min = (lower == up);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const unsigned long long ok = __shfl_xor_sync(FULL, K[r], d);
                        const unsigned oi = __shfl_xor_sync(FULL, I[r], d);
                        const bool other_less = (ok < K[r]) || (ok == K[r] && oi < I[r]);
                        const bool take = (other_less == keep_min);
                        K[r] = take ? ok : K[r];
                        I[r] = take ? oi : I[r];
                    }
                }
            }
        }
        unsigned long long kout[2];
        unsigned iout[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int e = 2 * lane + r;
            kout[r] = K[r];
            iout[r] = I[r];
            if (!use_key && e < m) {
                kout[r] = B.key[I[r]];
                iout[r] = (unsigned)K[r];
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int e = 2 * lane + r;
            if (e < m) {
                B.key[e] = kout[r];
                B.idx[e] = iout[r];
            }
        }
        __syncwarp();
        return;
    }
    for (int e = m + lane; e < p2; e += 32) {
        B.key[e] = ~0ull;
        B.idx[e] = 0xffffffffu;
    }
    __syncwarp();
    for (int size = 2; size <= p2; size <<= 1) {
        for (int ls = 31 - __clz(size >> 1); ls >= 0; --ls) {
            const int stride = 1 << ls;
            for (int tsk = lane; tsk < (p2 >> 1); tsk += 32) {
                const int i = ((tsk >> ls) << (ls + 1)) + (tsk & (stride - 1));
                const int j = i + stride;
                const bool up = ((i & size) == 0);
                const unsigned long long ki = B.key[i], kj = B.key[j];
                const unsigned ii = B.idx[i], ij = B.idx[j];
                const unsigned long long pi = use_key ? ki : (ii == 0xffffffffu ? ~0ull : 0ull);
                const unsigned long long pj = use_key ? kj : (ij == 0xffffffffu ? ~0ull : 0ull);
                const bool gt = (pi > pj) || (pi == pj && ii > ij);
                if (gt == up) {
                    B.key[i] = kj;
                    B.key[j] = ki;
                    B.idx[i] = ij;
                    B.idx[j] = ii;
                }
            }
            __syncwarp();
        }
    }
}


struct SearchParams {
    const double4 *obs;
    const unsigned *cell_start;
    const GridDims *gd;
    const double *tx, *ty, *tt;     // full target arrays
    const long long *leave_out;     // nullable
    const unsigned *order;          // chunk-local visit order
    int64_t c0;                     // first target of the chunk
    int n;                          // targets in the chunk
    double smax2, tmax;
    int k, min_pts;
    Model M;
    int *sel_n;            // [n]   -1: below min_pts (or no candidate), else neighbours kept
    unsigned *sel_idx;     // [n][k]
    double *sel_gamma;     // [n][k]
    unsigned *cand_n;      // [n] candidates inside the window
    unsigned *exam_n;      // [n] candidates examined
};

// gamma of the pending candidates -> (key, idx) buffer.  Unless `all`, only full groups of 32
// are evaluated (every lane busy) and the < 32 leftovers move to the front of the list.
template <int GV>
__device__ __forceinline__ void flush_pending(const WarpBuf &B, const SearchParams &P, int &pend_n,
                                              int &m, bool &truncated, bool all, int lane)
{
    __syncwarp();
    const int nproc = all ? pend_n : (pend_n & ~31);
    if (m + nproc > BUF_CAP) {
        select_k(B, m, P.k, lane);
        truncated = true;
    }
    for (int e = lane; e < nproc; e += 32) {
        const double h = sqrt(B.pend_d2[e]);              // predictor.py:627-634
        B.key[m + e] = gamma_key(gamma_search<GV>(P.M, h, B.pend_lt[e]));   // lag_t: predictor.py:635
        B.idx[m + e] = B.pend_idx[e];
    }
    m += nproc;
    const int rem = pend_n - nproc;
    if (rem > 0) {
        double d2 = 0.0, lt = 0.0;
        unsigned oi = 0u;
        if (lane < rem) {
            d2 = B.pend_d2[nproc + lane];
            lt = B.pend_lt[nproc + lane];
            oi = B.pend_idx[nproc + lane];
        }
        __syncwarp();
        if (lane < rem) {
            B.pend_d2[lane] = d2;
            B.pend_lt[lane] = lt;
            B.pend_idx[lane] = oi;
        }
    }
    pend_n = rem;
    __syncwarp();
}

template <int GV>
__global__ void __launch_bounds__(SEARCH_WARPS * 32) search_kernel(SearchParams P)
{
    __shared__ unsigned long long s_key[SEARCH_WARPS][BUF_CAP];
    __shared__ double s_pd2[SEARCH_WARPS][PEND_CAP];
    __shared__ double s_plt[SEARCH_WARPS][PEND_CAP];
    __shared__ unsigned s_idx[SEARCH_WARPS][BUF_CAP];
    __shared__ unsigned s_pidx[SEARCH_WARPS][PEND_CAP];
    __shared__ int s_rt[SEARCH_WARPS][32];          // run table: 16 virtual ends + 16 offsets
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * SEARCH_WARPS + warp;
    if (slot >= P.n) return;
    WarpBuf B{s_key[warp], s_idx[warp], s_pd2[warp], s_plt[warp], s_pidx[warp]};
    const unsigned tl = P.order[slot];
    const int64_t tg = P.c0 + tl;
    const double x = P.tx[tg], y = P.ty[tg], t = P.tt[tg];
    const long long leave = P.leave_out ? P.leave_out[tg] : -1;
    const GridDims g = *P.gd;

    const int cxr = raw_cell(x, g.x0, g.inv_cx, g.nx, g.hx);
    const int cyr = raw_cell(y, g.y0, g.inv_cy, g.ny, g.hy);
    const int ctr = raw_cell(t, g.t0, g.inv_ct, g.nt, g.ht);
    const int cx_lo = max(cxr - g.hx, 0), cx_hi = min(cxr + g.hx, g.nx - 1);
    const int cy_lo = max(cyr - g.hy, 0), cy_hi = min(cyr + g.hy, g.ny - 1);
    const int ct_lo = max(ctr - g.ht, 0), ct_hi = min(ctr, g.nt - 1);  // causal: t_obs <= t_tgt

    int pend_n = 0, m = 0;
    bool truncated = false;
    unsigned examined = 0, in_window = 0;
    const unsigned lt_mask = (1u << lane) - 1u;
    // The stencil is (cy, ct) rows of x-contiguous cells = contiguous runs of the cell-sorted
    // observations.  The runs are concatenated into one virtual sequence (run table in shared
    // memory, 16 runs at a time) so that every 32-candidate load is full.
    const int ncy = cy_hi - cy_lo + 1, nct = ct_hi - ct_lo + 1;
    const int nruns = (cx_lo <= cx_hi && ncy > 0 && nct > 0) ? ncy * nct : 0;
    int *rt_vend = s_rt[warp], *rt_off = s_rt[warp] + 16;
    for (int rbase = 0; rbase < nruns; rbase += 16) {
        unsigned beg = 0, len = 0;
        const int r = rbase + lane;
        if (lane < 16 && r < nruns) {
            const int ct = ct_lo + r / ncy, cy = cy_lo + r % ncy;
            const int row = (ct * g.ny + cy) * g.nx;
            beg = P.cell_start[row + cx_lo];
            len = P.cell_start[row + cx_hi + 1] - beg;
        }
        unsigned vend = len;                  // inclusive prefix sum of the run lengths
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
            const unsigned o = __shfl_up_sync(FULL, vend, off);
            if (lane >= off) vend += o;
        }
        const unsigned total = __shfl_sync(FULL, vend, 15);
        __syncwarp();
        if (lane < 16) {
            rt_vend[lane] = (int)vend;
            rt_off[lane] = (int)beg - (int)(vend - len);     // position = virtual index + off
        }
        __syncwarp();
        examined += total;
        for (unsigned v0 = 0; v0 < total; v0 += 32 * SCAN_U) {
            double4 o[SCAN_U];
#pragma unroll
            for (int u = 0; u < SCAN_U; ++u) {     // loads first: SCAN_U requests in flight
                const unsigned v = v0 + 32 * u + lane;
                o[u] = make_double4(0.0, 0.0, INFINITY, 0.0);     // t_obs = inf fails the filter
                if (v < total) {
                    int rr = 0;                    // first run with vend > v
                    if (rt_vend[rr + 7] <= (int)v) rr += 8;
                    if (rt_vend[rr + 3] <= (int)v) rr += 4;
                    if (rt_vend[rr + 1] <= (int)v) rr += 2;
                    if (rt_vend[rr] <= (int)v) rr += 1;
                    o[u] = P.obs[(int)v + rt_off[rr]];
                }
            }
#pragma unroll
            for (int u = 0; u < SCAN_U; ++u) {
                if (v0 + 32 * u < total) {         // warp-uniform
                    const double dx = o[u].x - x, dy = o[u].y - y;
                    const double d2 = dx * dx + dy * dy;      // cdist sqeuclidean, data.py:195-199
                    const double dt = o[u].z - t;
                    const unsigned oi = (unsigned)__double_as_longlong(o[u].w);
                    const bool pass = (d2 < P.smax2) && (fabs(dt) < P.tmax) && (o[u].z <= t) &&   // data.py:206-211
                                      ((long long)oi != leave);
                    const unsigned bal = __ballot_sync(FULL, pass);
                    if (pass) {
                        const int e = pend_n + __popc(bal & lt_mask);
                        B.pend_d2[e] = d2;
                        B.pend_lt[e] = fabs(dt);
                        B.pend_idx[e] = oi;
                    }
                    pend_n += __popc(bal);
                }
            }
            if (pend_n >= 32) {
                in_window += pend_n & ~31;
                flush_pending<GV>(B, P, pend_n, m, truncated, false, lane);
            }
        }
    }
    in_window += pend_n;
    flush_pending<GV>(B, P, pend_n, m, truncated, true, lane);
    if (m > P.k) {
        select_k(B, m, P.k, lane);
        truncated = true;
    }
    if (lane == 0) {
        P.cand_n[tl] = in_window;
        P.exam_n[tl] = examined;
    }
    if ((int)in_window < P.min_pts || in_window == 0) {  // predictor.py:624-626
        if (lane == 0) P.sel_n[tl] = ((int)in_window < P.min_pts) ? -1 : 0;
        return;
    }
    // reference order: ascending gamma when truncated (predictor.py:637-642), else the
    // pre-filter's ascending observation index (data.py:212)
    sort_buf(B, m, truncated, lane);
    for (int e = lane; e < m; e += 32) {
        P.sel_idx[(size_t)tl * P.k + e] = B.idx[e];
        P.sel_gamma[(size_t)tl * P.k + e] = key_gamma(B.key[e]);
    }
    if (lane == 0) P.sel_n[tl] = m;
}

// ---- assemble + solve + epilogue ---------------------------------------------------
struct SolveParams {
    const double *ox, *oy, *ot, *oresid;   // observations, original order
    const unsigned *order;
    int64_t c0;
    int n;
    int k, ld;
    int flags;
    Model M;
    const int *sel_n;
    const unsigned *sel_idx;
    const double *sel_gamma;
    double *mean, *std;          // [n_targets] (global target index)
    float *w_out, *g_out;        // nullable, [n_targets][k]
    unsigned *idx_out;           // nullable
    int *nused_out;              // nullable
    unsigned *singular;          // [n] chunk-local flag (bit 0 rank-deficient, bit 1 non-finite)
    unsigned *sing_list;         // rank-deficient targets (chunk-local ids) for the minimum-norm kernel
    unsigned *sing_count;
    const unsigned *count;       // nullable: number of entries of `order` to process (device)
    const double4 *obs4;         // observations {x, y, t, resid}, original order
    const unsigned *pairtab;     // pair e -> a | b << 8 | ... (b < a), see CholParams::pairtab
    // product / product_sum with spherical marginals and every neighbour pair inside both
    // ranges: gamma_s = pbs + h (pas - pcs h^2), gamma_t = pbt + t (pat - pct t^2)
    double pbs, pas, pcs, pbt, pat, pct;
};

enum { LU_GENERIC = 0, LU_PRODUCT_SPH = 1, LU_PRODUCT_SUM_SPH = 2 };



// numpy's pairwise sum of a row of length n <= 128 (numpy/_core/src/umath/loops_utils.h.src),
// preceded by the reduction's initial 0: bit-compatible with np.sum(row) -- checked against
// numpy 2.3 in tests/test_oracle_cpu.py.
template <typename T>
__device__ T numpy_row_sum(const T *a, int n)
{
    T res;
    if (n < 8) {
        res = (T)0;
        for (int i = 0; i < n; ++i) res = res + a[i];
    } else {
        T r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = r[j] + a[i + j];
        }
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res = res + a[i];
    }
    return (T)0 + res;
}

// Test hook (stif_debug_numpy_row_sum): one thread per row runs the device summation.
template <typename T>
__global__ void row_sum_kernel(const T *rows, int n_rows, int k, T *out)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) out[r] = numpy_row_sum<T>(rows + (size_t)r * k, k);
}

// ---- solve: shared pieces ------------------------------------------------------------
// One CTA of WPT warps per target.  The augmented system [A | b] ((m+1) x (m+2)) lives in
// shared memory, row-major with leading dimension ld (odd), zero-padded to NR x NC
// (multiples of 8) so the tensor-core tiles need no bounds checks.
struct SolveSmem {
    double *A;                       // NR x ld; column n = rhs
    double *xs, *ys, *ts, *sol, *g0; // [k+1] each
    unsigned *ids;                   // [k+1]
    int *piv;                        // [8] pivot rows of the current panel
    unsigned long long *scale;       // bits of max |entry| of the assembled system
    double *pvals;                   // [8] pivot row of the current column (panel entries), 16-B aligned
    unsigned long long *pkey;        // [8] per-warp pivot keys of the current column
};

__device__ __forceinline__ SolveSmem carve(double *base, int k, int ld, int nr)
{
    SolveSmem S;
    S.A = base;
    S.pvals = base + (size_t)nr * ld;            // nr is a multiple of 8: 16-byte aligned
    S.pkey = reinterpret_cast<unsigned long long *>(S.pvals + 8);
    S.xs = S.pvals + 16;
    S.ys = S.xs + (k + 1);
    S.ts = S.ys + (k + 1);
    S.sol = S.ts + (k + 1);
    S.g0 = S.sol + (k + 1);
    S.scale = reinterpret_cast<unsigned long long *>(S.g0 + (k + 1));
    S.piv = reinterpret_cast<int *>(S.scale + 1);
    S.ids = reinterpret_cast<unsigned *>(S.piv + 8);
    return S;
}

// Packed pivot key: |value| bits with the low byte replaced by (255 - row); one unsigned max
// yields the largest magnitude and, among (near-)equal ones, the smallest row.  The warp max
// uses two 32-bit REDUX instead of a 5-round 64-bit shuffle tree.
__device__ __forceinline__ unsigned long long pivot_key(double v, int r)
{
    return ((unsigned long long)__double_as_longlong(fabs(v)) & ~0xffull) | (unsigned long long)(255 - r);
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long key)
{
    const unsigned hi = (unsigned)(key >> 32);
    const unsigned mh = __reduce_max_sync(FULL, hi);
    const unsigned lo = (hi == mh) ? (unsigned)key : 0u;
    const unsigned ml = __reduce_max_sync(FULL, lo);
    return ((unsigned long long)mh << 32) | ml;
}

// Loads the selection, assembles [A | b] (predictor.py:573-585) and zero-fills the padding.
// Returns m (neighbours used); m <= 0 means the outputs were written (zeros) and the CTA is done.
// One neighbour per thread, fetched ahead of time into registers by the persistent loop.
struct LuPrefetch {
    unsigned tl;      // chunk-local target id
    int m;            // neighbours (<= 0: nothing to solve)
    unsigned id;      // this thread's neighbour (e = tid)
    double4 o;        // its {x, y, t, resid}
    double g;         // its gamma-vector entry
};

// Loads the selection, assembles [A | b] (predictor.py:573-585) and zero-fills the padding.
// Returns m (neighbours used); m <= 0 means the outputs were written (zeros) and the CTA is done.
template <int NT, int LUV>
__device__ int solve_setup(const SolveParams &P, const SolveSmem &S, const LuPrefetch &F, int64_t tg,
                           int tid, int nr, int nc, bool &nonfinite)
{
    nonfinite = false;
    const int k = P.k, ld = P.ld;
    double *A = S.A;
    const int m = F.m;
    if (m <= 0) {
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
        return m;
    }
    const int n = m + 1;
    // padding: rows n..nr-1 (all columns) and columns n+1..nc-1 of the real rows
    for (int e = n * ld + tid; e < nr * ld; e += NT) A[e] = 0.0;
    for (int r = tid; r < n; r += NT)
        for (int c = n + 1; c < nc; ++c) A[r * ld + c] = 0.0;
    if (tid < m) {                // NT >= k + 1: one neighbour per thread
        const int e = tid;
        S.ids[e] = F.id;
        S.xs[e] = F.o.x;
        S.ys[e] = F.o.y;
        S.ts[e] = F.o.z;
        S.g0[e] = F.g;
        A[e * ld + n] = F.g;      // rhs = [gamma_vector; 1]   predictor.py:582-585
        A[e * ld + m] = 1.0;      // Lagrange column            predictor.py:578
        A[m * ld + e] = 1.0;      // Lagrange row
    }
    if (tid == 0) {
        A[m * ld + m] = 0.0;      // predictor.py:580
        A[m * ld + n] = 1.0;
        *S.scale = 0ull;
    }
    __syncthreads();
    // gamma matrix, predictor.py:573-575 (diagonal = gamma(0,0), no special case)
    const double gdiag = gamma_st(P.M, 0.0, 0.0);
    double amax = fmax(fabs(gdiag), 1.0);   // largest |entry| (the border holds ones)
    // NaN / Inf anywhere in [A | b]: the reference's lstsq raises LinAlgError (numba
    // _check_finite_matrix, predictor.py:587); flagged here, reported by the host entry point
    bool bad = !(fabs(gdiag) < INFINITY) || (tid < m && !(fabs(F.g) < INFINITY));
    for (int e = tid; e < m; e += NT) A[e * ld + e] = gdiag;
    const int npairs = m * (m - 1) / 2;
    for (int e = tid; e < npairs; e += NT) {
        const unsigned w = __ldg(P.pairtab + e);
        const int a = w & 0xff, b = (w >> 8) & 0xff;       // b < a
        const double dx = S.xs[a] - S.xs[b], dy = S.ys[a] - S.ys[b];
        const double lt = fabs(S.ts[a] - S.ts[b]);         // utils.py:19-22
        double g;
        if (LUV == LU_GENERIC) {
            const double h = sqrt(dx * dx + dy * dy);      // utils.py:40-44
            g = gamma_st(P.M, h, lt);
        } else {
            const double d2 = __fma_rn(dy, dy, dx * dx) + 1e-290;    // sqrt(0) -> 1e-145, no 0 * inf
            const double h = sqrt_cubic_lu(d2);
            const double gs = __fma_rn(h, __fma_rn(-P.pcs, d2, P.pas), P.pbs);
            const double gt = __fma_rn(lt, __fma_rn(-P.pct, lt * lt, P.pat), P.pbt);
            const double pr = gs * gt;
            g = (LUV == LU_PRODUCT_SUM_SPH) ? __fma_rn(P.M.k, pr, pr) : pr;
        }
        A[a * ld + b] = g;
        A[b * ld + a] = g;
        amax = fmax(amax, fabs(g));         // fmax ignores NaN
        bad = bad || !(fabs(g) < INFINITY);
    }
    {
        const unsigned long long wm = warp_max_u64((unsigned long long)__double_as_longlong(amax));
        if ((tid & 31) == 0) atomicMax(S.scale, wm);
    }
    nonfinite = __syncthreads_or(bad ? 1 : 0) != 0;
    return m;
}

// Fused epilogue (one warp): predictor.py:646-658 (storage) and :900-905 (mean, std).
// sol/g0/ids: the m weights, gamma-vector entries and observation indices; rloc: residuals of
// the neighbours if the caller staged them (else gathered from P.oresid); sc64/sc32a/sc32b:
// scratch for k doubles / k floats / k floats.
__device__ void solve_epilogue(const SolveParams &P, const double *sol, const double *g0,
                               const unsigned *ids, const double *rloc, double *sc64, float *sc32a,
                               float *sc32b, int m, int64_t tg, int lane)
{
    const int k = P.k;
    if (P.flags & STIF_KRIG_F32_STORAGE) {
        double *term64 = sc64;
        float *w32 = sc32a;
        float *term32 = sc32b;
        const double resid0 = P.oresid[0];
        __syncwarp();
        for (int e = lane; e < k; e += 32) {
            float w = 0.f, g = 0.f;
            unsigned id = 0u;
            double r = resid0;                            // padded slots index observation 0
            if (e < m) {
                w = (float)sol[e];
                g = (float)g0[e];
                id = ids[e];
                r = rloc ? rloc[e] : P.oresid[id];
            }
            w32[e] = w;
            term64[e] = (double)w * r;                    // f32 * f64 -> f64
            term32[e] = w * g;                            // f32 * f32 -> f32
            if (P.w_out) P.w_out[(size_t)tg * k + e] = w;
            if (P.g_out) P.g_out[(size_t)tg * k + e] = g;
            if (P.idx_out) P.idx_out[(size_t)tg * k + e] = id;
        }
        __syncwarp();
        if (lane == 0) {
            P.mean[tg] = numpy_row_sum<double>(term64, k);
            P.std[tg] = (double)sqrtf(numpy_row_sum<float>(term32, k));
            if (P.nused_out) P.nused_out[tg] = m;
        }
    } else {
        double am = 0.0, av = 0.0;
        for (int e = lane; e < m; e += 32) {
            am += sol[e] * (rloc ? rloc[e] : P.oresid[ids[e]]);
            av += sol[e] * g0[e];
        }
        for (int off = 16; off > 0; off >>= 1) {
            am += __shfl_xor_sync(FULL, am, off);
            av += __shfl_xor_sync(FULL, av, off);
        }
        for (int e = lane; e < k; e += 32) {
            const bool in = e < m;
            if (P.w_out) P.w_out[(size_t)tg * k + e] = in ? (float)sol[e] : 0.f;
            if (P.g_out) P.g_out[(size_t)tg * k + e] = in ? (float)g0[e] : 0.f;
            if (P.idx_out) P.idx_out[(size_t)tg * k + e] = in ? ids[e] : 0u;
        }
        if (lane == 0) {
            P.mean[tg] = am;
            P.std[tg] = sqrt(av);
            if (P.nused_out) P.nused_out[tg] = m;
        }
    }
    __syncwarp();
}

// Back substitution of the LU (warp 0; column oriented: x_j = b_j * (1/U_jj), then
// b_r -= U_rj x_j for r < j), then the epilogue.
__device__ void solve_finish(const SolveParams &P, const SolveSmem &S, int m, unsigned tl, int64_t tg,
                             int lane, bool singular, bool nonfinite)
{
    const int ld = P.ld, n = m + 1;
    double *A = S.A, *sol = S.sol;
    for (int j = n - 1; j >= 0; --j) {
        const double xj = A[j * ld + n] * A[j * ld + j];
        for (int r = lane; r < j; r += 32) A[r * ld + n] = __fma_rn(-A[r * ld + j], xj, A[r * ld + n]);
        if (lane == 0) sol[j] = xj;
        __syncwarp();
    }
    // bit 0: rank-deficient (a pivot below the tolerance); bit 1: NaN / Inf in the system
    if (lane == 0 && (singular || nonfinite)) {
        P.singular[tl] = nonfinite ? 2u : 1u;
        if (!nonfinite) P.sing_list[atomicAdd(P.sing_count, 1u)] = tl;
    }
    solve_epilogue(P, sol, S.g0, S.ids, nullptr, S.xs, reinterpret_cast<float *>(S.ys),
                   reinterpret_cast<float *>(S.ts), m, tg, lane);
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- solve: blocked LU, trailing update on the FP64 tensor cores -----------------------
// Right-looking blocked LU with partial pivoting, block 8.  Per panel:
//   1. warp 0 factorises the 8-column panel (lanes over rows; REDUX pivot search; row swaps
//      over all columns from the panel start; NEGATED multipliers stored in place; 1/pivot on
//      the diagonal, 0 = column skipped because it is exactly singular or NaN);
//   2. U12 = L11^-1 A12: one thread per column to the right of the panel;
//   3. A22 += (-L21) U12 as 8x8 tiles of two mma.sync.m8n8k4.f64 (DMMA) each, tiles dealt
//      round-robin to the warps.  512 FMAs per tensor instruction instead of 32 per DFMA: the
//      scalar shared-memory LU was issue-bound (ncu: 78 % issue-active at 27 % FP64 pipe).
// The right-hand side is column n and is transformed with the matrix.
template <int WPT, int LUV>
__global__ void __launch_bounds__(WPT * 32) solve_blk_kernel(SolveParams P, int nr, int nc)
{
    extern __shared__ __align__(16) double smem_d[];
    constexpr int NT = WPT * 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ld = P.ld;
    const SolveSmem S = carve(smem_d, P.k, ld, nr);
    double *A = S.A;
    const int n_work = P.count ? (int)*P.count : P.n;
    // persistent CTAs: targets in visit order, strided over the grid.  The next target's
    // selection is fetched into registers (one neighbour per thread) in three stages spread
    // over the current target's phases, so no dependent global-load chain sits at the loop top.
    LuPrefetch F, N;
    F.tl = 0u; F.m = 0; F.id = 0u; F.o = make_double4(0.0, 0.0, 0.0, 0.0); F.g = 0.0;
    N = F;
    bool has_next = false;
    auto fetch_ids = [&]() {
        N.m = has_next ? P.sel_n[N.tl] : 0;
        if (has_next && tid < P.k) N.id = P.sel_idx[(size_t)N.tl * P.k + tid];
    };
    auto fetch_obs = [&]() {
        if (tid < N.m) {
            N.o = P.obs4[N.id];
            N.g = P.sel_gamma[(size_t)N.tl * P.k + tid];
        }
    };
    if ((int)blockIdx.x < n_work) {
        has_next = true;
        N.tl = P.order[blockIdx.x];
        fetch_ids();
        fetch_obs();
    }
    for (int slot = blockIdx.x; slot < n_work; slot += gridDim.x) {
        __syncthreads();   // the previous target's epilogue is done with the shared arrays
        F = N;
        const unsigned tl = F.tl;
        const int64_t tg = P.c0 + tl;
        has_next = slot + (int)gridDim.x < n_work;
        if (has_next) N.tl = P.order[slot + gridDim.x];
        bool nonfinite = false;
        const int m = solve_setup<NT, LUV>(P, S, F, tg, tid, nr, nc, nonfinite);
        fetch_ids();
        if (m <= 0) {
            fetch_obs();
            continue;
        }
        const int n = m + 1;
        bool singular = false;
        // rank tolerance: a pivot below n * 8 eps * max|entry| is rounding noise of an exactly
        // singular system (duplicate observations); that column is skipped (x_j := 0).  dgelsd
        // (rcond = eps) would return the minimum-norm solution instead: best-effort parity.
        const double tol = __longlong_as_double((long long)*S.scale) * (1.8e-15 * (double)n);

        for (int jb = 0; jb < n; jb += 8) {
            {
                // ---- 1. panel factorisation, columns jb .. jb+7 (pivot columns only while < n) ----
                // The whole CTA: thread r owns row r and keeps its 8 panel entries in registers.
                // Rows are not moved: `pos` is the row's position after the swaps so far.  Per
                // column: warp-level REDUX on a packed magnitude/row key, the per-warp keys meet
                // in shared memory (barrier), the owner of the pivot row publishes its entries
                // (barrier), every thread eliminates its own row.  Rows are written back to their
                // final positions at the end.
                double a[8];
                int pos = tid;
#pragma unroll
                for (int c = 0; c < 8; ++c) a[c] = (tid >= jb && tid < n) ? A[tid * ld + jb + c] : 0.0;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = jb + jj;
                    if (j < n) {          // CTA-uniform
                        unsigned long long key = 0ull;
                        if (pos >= j && pos < n) key = pivot_key(a[jj], pos);
                        key = warp_max_u64(key);
                        if (lane == 0) S.pkey[warp] = key;
                        __syncthreads();
#pragma unroll
                        for (int w = 0; w < WPT; ++w) {
                            const unsigned long long kw = S.pkey[w];
                            key = kw > key ? kw : key;
                        }
                        const int bi = 255 - (int)(key & 0xffull);
                        const double bv = __longlong_as_double((long long)(key & ~0xffull));
                        const bool ok = bv > tol && bv < INFINITY;
                        if (!ok) singular = true;
                        const int prow = ok ? bi : j;     // row swapped with row j (j = no swap)
                        if (pos == prow) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) S.pvals[c] = a[c];
                        }
                        __syncthreads();
                        double u[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) u[c] = (c >= jj) ? S.pvals[c] : 0.0;
                        if (pos == prow) pos = j;
                        else if (pos == j) pos = prow;
                        double pinv = 0.0;
                        if (ok) {      // 1 / pivot: MUFU.RCP64H + two Newton steps
                            double y;
                            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u[jj]));
                            y = __fma_rn(y, __fma_rn(-u[jj], y, 1.0), y);
                            pinv = __fma_rn(y, __fma_rn(-u[jj], y, 1.0), y);
                        }
                        if (tid == 0) S.piv[jj] = prow;
                        if (pos > j && pos < n) {
                            const double mneg = -(a[jj] * pinv);   // negated multiplier
                            a[jj] = mneg;
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                if (c > jj) a[c] = __fma_rn(mneg, u[c], a[c]);
                        } else if (pos == j) {
                            a[jj] = pinv;
                        }
                    }
                }
                if (tid >= jb && tid < n) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) A[pos * ld + jb + c] = a[c];
                }
            }
            __syncthreads();
            if (jb + 8 > n) break;    // rhs column inside the panel: nothing left to transform
            // ---- 2. U12: rows jb..jb+7 of the columns right of the panel ----
            for (int c = jb + 8 + tid; c < nc; c += NT) {
                // deferred row swaps of this panel, in pivot order
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int pr = S.piv[i];
                    if (pr != jb + i) {
                        const double tmp = A[(jb + i) * ld + c];
                        A[(jb + i) * ld + c] = A[pr * ld + c];
                        A[pr * ld + c] = tmp;
                    }
                }
                double uu[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    double acc = A[(jb + i) * ld + c];
#pragma unroll
                    for (int q = 0; q < i; ++q) acc = __fma_rn(A[(jb + i) * ld + jb + q], uu[q], acc);
                    uu[i] = acc;
                }
#pragma unroll
                for (int i = 1; i < 8; ++i) A[(jb + i) * ld + c] = uu[i];
            }
            __syncthreads();
            // ---- 3. trailing update on the tensor cores ----
            {
                const int g = lane >> 2, t = lane & 3;
                const int r_lo = jb + 8, ntr = (nr - r_lo) >> 3, ntc = (nc - r_lo) >> 3;
                for (int tile = warp; tile < ntr * ntc; tile += WPT) {
                    const int ti = tile / ntc, tj = tile - ti * ntc;
                    const int r0 = r_lo + 8 * ti, c0 = r_lo + 8 * tj;
                    const double a0 = A[(r0 + g) * ld + jb + t], a1 = A[(r0 + g) * ld + jb + 4 + t];
                    const double b0 = A[(jb + t) * ld + c0 + g], b1 = A[(jb + 4 + t) * ld + c0 + g];
                    double *cp = A + (r0 + g) * ld + c0 + 2 * t;
                    double c0v = cp[0], c1v = cp[1];
                    dmma884(c0v, c1v, a0, b0);
                    dmma884(c0v, c1v, a1, b1);
                    cp[0] = c0v;
                    cp[1] = c1v;
                }
            }
            __syncthreads();
        }
        fetch_obs();
        if (warp == 0) solve_finish(P, S, m, tl, tg, lane, singular, nonfinite);
    }
}

// ---- minimum-norm solve of rank-deficient systems ------------------------------------------
// The reference solves every system with np.linalg.lstsq = LAPACK dgelsd (SVD, rcond = eps),
// which returns the MINIMUM-NORM solution when [Gamma 1; 1' 0] is rank-deficient -- duplicate or
// co-located observations make two rows equal whatever the nugget, because the reference puts
// gamma(0, 0) on the diagonal (stif/predictor.py:573-587).  Targets the pivoted LU flags (a pivot
// below n * 8 eps * max|entry|) are re-solved here the same way: one-sided Jacobi (Hestenes) SVD
// of the symmetric matrix, G = A V with orthogonal columns g_j = sigma_j u_j, then
//      x = sum over sigma_j > thr * sigma_max of  v_j (g_j . b) / sigma_j^2.
// One CTA per target; G lives in shared memory (column-major), V in a global scratch slab.
// A rare path (statistics out[3] counts it): clarity over speed.
constexpr int MN_THREADS = 128;

__global__ void __launch_bounds__(MN_THREADS) solve_minnorm_kernel(SolveParams P, double *vscratch,
                                                                   double *w64_out /* nullable: float64 weights */)
{
    extern __shared__ __align__(16) double smem_d[];
    constexpr int NW = MN_THREADS / 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = P.k, nmax = k + 1;
    double *G = smem_d;                       // [n][n] column-major
    double *xs = G + (size_t)nmax * nmax, *ys = xs + nmax, *ts = ys + nmax, *rs = ts + nmax;
    double *g0 = rs + nmax, *bv = g0 + nmax, *sol = bv + nmax, *coef = sol + nmax, *alpha = coef + nmax;
    double *sc64 = alpha + nmax;
    float *sc32 = reinterpret_cast<float *>(sc64 + nmax);
    unsigned *ids = reinterpret_cast<unsigned *>(sc32 + 2 * nmax);
    double *V = vscratch + (size_t)blockIdx.x * nmax * nmax;
    const int n_work = (int)*P.sing_count;
    const double gdiag = gamma_st(P.M, 0.0, 0.0);
    for (int slot = blockIdx.x; slot < n_work; slot += gridDim.x) {
        __syncthreads();
        const unsigned tl = P.sing_list[slot];
        const int64_t tg = P.c0 + tl;
        const int m = P.sel_n[tl];
        if (m <= 0) continue;
        const int n = m + 1;
        for (int e = tid; e < m; e += MN_THREADS) {
            const unsigned id = P.sel_idx[(size_t)tl * k + e];
            const double4 o = P.obs4[id];
            ids[e] = id;
            xs[e] = o.x;
            ys[e] = o.y;
            ts[e] = o.z;
            rs[e] = o.w;
            g0[e] = P.sel_gamma[(size_t)tl * k + e];
            bv[e] = g0[e];
        }
        if (tid == 0) bv[m] = 1.0;
        __syncthreads();
        // the reference's matrix verbatim (predictor.py:573-580, utils.py:19-22, :40-44)
        for (int e = tid; e < n * n; e += MN_THREADS) {
            const int c = e / n, r = e - c * n;
            double val;
            if (r == m || c == m) val = (r == c) ? 0.0 : 1.0;
            else if (r == c) val = gdiag;
            else {
                const double dx = xs[r] - xs[c], dy = ys[r] - ys[c];
                val = gamma_st(P.M, sqrt(dx * dx + dy * dy), fabs(ts[r] - ts[c]));
            }
            G[e] = val;
            V[e] = (r == c) ? 1.0 : 0.0;
        }
        __syncthreads();
        double fro2 = 0.0;
        for (int e = tid; e < n * n; e += MN_THREADS) fro2 = __fma_rn(G[e], G[e], fro2);
        for (int off = 16; off > 0; off >>= 1) fro2 += __shfl_xor_sync(FULL, fro2, off);
        if (lane == 0) alpha[warp] = fro2;
        __syncthreads();
        fro2 = 0.0;
        for (int w = 0; w < NW; ++w) fro2 += alpha[w];
        __syncthreads();
        const int ne = n + (n & 1), half = ne >> 1;
        for (int sweep = 0; sweep < 60; ++sweep) {
            bool rotated = false;
            for (int round = 0; round < ne - 1; ++round) {
                for (int i = warp; i < half; i += NW) {      // round-robin tournament: disjoint pairs
                    int p = (i == 0) ? ne - 1 : (round + i) % (ne - 1);
                    int q = (i == 0) ? round : (round + (ne - 1) - i) % (ne - 1);
                    if (p >= n || q >= n) continue;
                    if (p > q) { const int tmp = p; p = q; q = tmp; }
                    double *gp = G + (size_t)p * n, *gq = G + (size_t)q * n;
                    double a = 0.0, b = 0.0, g = 0.0;
                    for (int r = lane; r < n; r += 32) {
                        const double u = gp[r], v = gq[r];
                        a = __fma_rn(u, u, a);
                        b = __fma_rn(v, v, b);
                        g = __fma_rn(u, v, g);
                    }
                    for (int off = 16; off > 0; off >>= 1) {
                        a += __shfl_xor_sync(FULL, a, off);
                        b += __shfl_xor_sync(FULL, b, off);
                        g += __shfl_xor_sync(FULL, g, off);
                    }
                    if (!(fabs(g) > 8e-15 * sqrt(a * b)) || !(fabs(g) > 1e-31 * fro2)) continue;
                    rotated = true;
                    const double zeta = (b - a) / (2.0 * g);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                    double *vp = V + (size_t)p * n, *vq = V + (size_t)q * n;
                    for (int r = lane; r < n; r += 32) {
                        const double u = gp[r], v = gq[r];
                        gp[r] = cs * u - sn * v;
                        gq[r] = sn * u + cs * v;
                        const double vu = vp[r], vv = vq[r];
                        vp[r] = cs * vu - sn * vv;
                        vq[r] = sn * vu + cs * vv;
                    }
                }
                __syncthreads();
            }
            if (!__syncthreads_or(rotated ? 1 : 0)) break;
        }
        // sigma_j^2 and g_j . b
        for (int j = warp; j < n; j += NW) {
            const double *gj = G + (size_t)j * n;
            double a = 0.0, d = 0.0;
            for (int r = lane; r < n; r += 32) {
                a = __fma_rn(gj[r], gj[r], a);
                d = __fma_rn(gj[r], bv[r], d);
            }
            for (int off = 16; off > 0; off >>= 1) {
                a += __shfl_xor_sync(FULL, a, off);
                d += __shfl_xor_sync(FULL, d, off);
            }
            if (lane == 0) {
                alpha[j] = a;
                coef[j] = d;
            }
        }
        __syncthreads();
        double amax = 0.0;
        for (int j = 0; j < n; ++j) amax = fmax(amax, alpha[j]);
        const double thr = 1.8e-15 * (double)n;               // the LU's rank tolerance
        const double cut = amax * thr * thr;
        __syncthreads();
        for (int j = tid; j < n; j += MN_THREADS) coef[j] = (alpha[j] > cut) ? coef[j] / alpha[j] : 0.0;
        __syncthreads();
        for (int r = tid; r < m; r += MN_THREADS) {
            double acc = 0.0;
            for (int j = 0; j < n; ++j) acc = __fma_rn(V[(size_t)j * n + r], coef[j], acc);
            sol[r] = acc;
        }
        __syncthreads();
        if (w64_out)
            for (int r = tid; r < m; r += MN_THREADS) w64_out[(size_t)tg * k + r] = sol[r];
        if (warp == 0)
            solve_epilogue(P, sol, g0, ids, rs, sc64, sc32, sc32 + nmax, m, tg, lane);
    }
}

#include "krig_chol.cuh"
#include "krig_ldl.cuh"

// stats[0] += sum(exam_n), [1] += sum(cand_n), [2] += #(sel_n < 0), [3] += #(rank-deficient),
// [4] += targets the covariance-form kernel handed to the pivoted LU, [5] += #(non-finite systems)
__global__ void stats_kernel(const unsigned *exam_n, const unsigned *cand_n, const int *sel_n,
                             const unsigned *singular, int n, const unsigned *fail_count,
                             unsigned long long *stats)
{
    if (fail_count && blockIdx.x == 0 && threadIdx.x == 0 && *fail_count)
        atomicAdd(&stats[4], (unsigned long long)*fail_count);
    unsigned long long a = 0, b = 0, c = 0, d = 0, f = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        a += exam_n[i];
        b += cand_n[i];
        c += sel_n[i] < 0 ? 1 : 0;
        d += singular[i] & 1u;
        f += (singular[i] >> 1) & 1u;
    }
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(FULL, a, off);
        b += __shfl_xor_sync(FULL, b, off);
        c += __shfl_xor_sync(FULL, c, off);
        d += __shfl_xor_sync(FULL, d, off);
        f += __shfl_xor_sync(FULL, f, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (a) atomicAdd(&stats[0], a);
        if (b) atomicAdd(&stats[1], b);
        if (c) atomicAdd(&stats[2], c);
        if (d) atomicAdd(&stats[3], d);
        if (f) atomicAdd(&stats[5], f);
    }
}

static int build_model(const KrigModel &km, Model &M)
{
    memset(&M, 0, sizeof(M));
    M.st = km.st_model;
    M.ks = km.space_model;
    M.kt = km.time_model;
    M.km = km.metric_model;
    const double *x = km.x;
    auto inv_of = [](int kind, double r) {
        if (kind == STIF_M_SPHERICAL) return 1.0 / (r / 1.0);   // a = r/1.   variogram_models.py:40
        const double a = r / 2.0;                                // a = r/2.   variogram_models.py:70
        return 1.0 / (a * a);
    };
    switch (km.st_model) {
    case STIF_ST_SUM:
    case STIF_ST_PRODUCT:
    case STIF_ST_PRODUCT_SUM:  // x = [r_s, r_t, c0_s, c0_t, b_s, b_t (, k)]
        M.rs = x[0]; M.rt = x[1]; M.c0s = x[2]; M.c0t = x[3]; M.bs = x[4]; M.bt = x[5];
        M.k = (km.st_model == STIF_ST_PRODUCT_SUM) ? x[6] : 0.0;
        break;
    case STIF_ST_METRIC:       // x = [r, c0, b, ani]
        M.rm = x[0]; M.c0m = x[1]; M.bm = x[2]; M.ani2 = x[3] * x[3];
        break;
    case STIF_ST_SUM_METRIC:   // x = [r_s, r_t, r_m, c0_s, c0_t, c0_m, b_s, b_t, b_m, ani]
        M.rs = x[0]; M.rt = x[1]; M.rm = x[2]; M.c0s = x[3]; M.c0t = x[4]; M.c0m = x[5];
        M.bs = x[6]; M.bt = x[7]; M.bm = x[8]; M.ani2 = x[9] * x[9];
        break;
    default:
        set_error("unknown space-time model id %d", km.st_model);
        return STIF_EINVAL;
    }
    M.is = inv_of(M.ks, M.rs);
    M.it = inv_of(M.kt, M.rt);
    M.im = inv_of(M.km, M.rm);
    return STIF_OK;
}

// ---- stand-alone neighbour pair list (Data.get_kriging_idxs, stif/data.py:170-212) --------
struct PairsParams {
    const double4 *obs;
    const unsigned *cell_start;
    const GridDims *gd;
    const double *tx, *ty, *tt;
    const long long *leave_out;        // nullable
    int n;
    double smax2, tmax;
    unsigned long long *counts;        // [n] pairs per target (count pass)
    const unsigned long long *offsets; // [n] exclusive prefix (fill pass)
    unsigned long long *keys;          // (obs_idx << 32) | target (fill pass)
};

template <bool FILL>
__global__ void __launch_bounds__(128) pairs_kernel(PairsParams P)
{
    const int lane = threadIdx.x & 31;
    const int tg = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (tg >= P.n) return;
    const double x = P.tx[tg], y = P.ty[tg], t = P.tt[tg];
    const long long leave = P.leave_out ? P.leave_out[tg] : -1;
    const GridDims g = *P.gd;
    const int cxr = raw_cell(x, g.x0, g.inv_cx, g.nx, g.hx);
    const int cyr = raw_cell(y, g.y0, g.inv_cy, g.ny, g.hy);
    const int ctr = raw_cell(t, g.t0, g.inv_ct, g.nt, g.ht);
    const int cx_lo = max(cxr - g.hx, 0), cx_hi = min(cxr + g.hx, g.nx - 1);
    const int cy_lo = max(cyr - g.hy, 0), cy_hi = min(cyr + g.hy, g.ny - 1);
    const int ct_lo = max(ctr - g.ht, 0), ct_hi = min(ctr, g.nt - 1);
    unsigned long long cnt = 0;
    const unsigned long long base = FILL ? P.offsets[tg] : 0ull;
    if (cx_lo <= cx_hi) {
        for (int ct = ct_lo; ct <= ct_hi; ++ct) {
            for (int cy = cy_lo; cy <= cy_hi; ++cy) {
                const int row = (ct * g.ny + cy) * g.nx;
                const unsigned beg = P.cell_start[row + cx_lo], end = P.cell_start[row + cx_hi + 1];
                for (unsigned p0 = beg; p0 < end; p0 += 32) {
                    const unsigned p = p0 + lane;
                    bool pass = false;
                    unsigned oi = 0u;
                    if (p < end) {
                        const double4 o = P.obs[p];
                        const double dx = o.x - x, dy = o.y - y;
                        const double d2 = dx * dx + dy * dy;            // data.py:195-199
                        oi = (unsigned)__double_as_longlong(o.w);
                        pass = (d2 < P.smax2) && (fabs(o.z - t) < P.tmax) && (o.z <= t) &&   // data.py:206-211
                               ((long long)oi != leave);
                    }
                    const unsigned bal = __ballot_sync(FULL, pass);
                    if (FILL && pass)
                        P.keys[base + cnt + __popc(bal & ((1u << lane) - 1u))] =
                            ((unsigned long long)oi << 32) | (unsigned)tg;
                    cnt += __popc(bal);
                }
            }
        }
    }
    if (!FILL && lane == 0) P.counts[tg] = cnt;
}

__global__ void pairs_unpack_kernel(const unsigned long long *keys, int64_t n, long long *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        out[2 * i] = (long long)(keys[i] >> 32);
        out[2 * i + 1] = (long long)(keys[i] & 0xffffffffull);
    }
}

// Weighted mean square error of the model on the lag grid (variogram_models.py:264-308).
// One block; per-thread partial sums in row-major order, fixed shared-memory tree.
__global__ void __launch_bounds__(256) wmse_kernel(Model M, const double *bins_space, int ns,
                                                   const double *bins_time, int nt, const double *emp,
                                                   const double *w, double *out)
{
    __shared__ double s_num[256], s_den[256];
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < ns * nt; i += 256) {
        const double g = gamma_st(M, bins_space[i / nt], bins_time[i % nt]);
        const double e = emp[i] - g;
        num += w[i] * (e * e);
        den += w[i];
    }
    s_num[threadIdx.x] = num;
    s_den[threadIdx.x] = den;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            s_num[threadIdx.x] += s_num[threadIdx.x + off];
            s_den[threadIdx.x] += s_den[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s_num[0] / s_den[0];
}

__global__ void obs_resid_kernel(const double *r, int64_t n, double4 *obs4)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) obs4[i].w = r[i];
}

__global__ void obs_interleave_kernel(const double *x, const double *y, const double *t,
                                      const double *r, int64_t n, double4 *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double4(x[i], y[i], t[i], r[i]);
}

// Host: pair table of the assembly loop (see CholParams::pairtab) for up to KMAX neighbours.
static void build_pair_table(std::vector<unsigned> &tab)
{
    tab.resize((size_t)KMAX * (KMAX - 1) / 2);
    size_t e = 0;
    for (int a = 1; a < KMAX; ++a) {
        for (int b = 0; b < a; ++b) {
            const int I = a >> 3, ri = a & 7, ci = b & 7;
            const int off = (I * (I + 1) / 2 + (b >> 3)) * 64 + ri * 8 +
                            ((((ci >> 1) ^ chol_swz(ri)) << 1) | (ci & 1));
            tab[e++] = (unsigned)a | ((unsigned)b << 8) | ((unsigned)off << 16);
        }
    }
}

// Host: decide whether the covariance-form (Cholesky) kernel may be tried for this model and
// fill its constants.  Returns the kernel variant (CHOL_*) or -1 for "pivoted LU only".
static int chol_plan(const KrigModel &km, const Model &M, int flags, double smax, double tmax,
                     CholParams &X)
{
    memset(&X, 0, sizeof(X));
    if (flags & STIF_KRIG_FORCE_LU) return -1;
    const bool use_s = km.st_model != STIF_ST_METRIC;
    const bool use_t = use_s;
    const bool use_m = km.st_model == STIF_ST_METRIC || km.st_model == STIF_ST_SUM_METRIC;
    // A sum of marginals is a valid variogram iff every partial sill c0 is >= 0, every range is
    // positive and the TOTAL nugget is >= 0: the marginal nuggets only enter through their sum
    // (unconstrained Nelder-Mead fits often return large nuggets of opposite sign).
    auto ok = [](double r, double c0, double b) {
        return r > 0.0 && isfinite(r) && c0 >= 0.0 && isfinite(c0) && isfinite(b);
    };
    bool valid = (km.st_model == STIF_ST_SUM || km.st_model == STIF_ST_METRIC ||
                  km.st_model == STIF_ST_SUM_METRIC);
    bool plain = true;      // no marginal nugget is negative: nothing cancels in the sill
    if (use_s) { valid = valid && ok(M.rs, M.c0s, M.bs); plain = plain && M.bs >= 0.0; }
    if (use_t) { valid = valid && ok(M.rt, M.c0t, M.bt); plain = plain && M.bt >= 0.0; }
    if (use_m) {
        valid = valid && ok(M.rm, M.c0m, M.bm) && M.ani2 >= 0.0 && isfinite(M.ani2);
        plain = plain && M.bm >= 0.0;
    }
    {
        const double nug_total = (use_s ? M.bs : 0.0) + (use_t ? M.bt : 0.0) + (use_m ? M.bm : 0.0);
        valid = valid && nug_total >= 0.0;
    }
    if (!valid && !(flags & STIF_KRIG_TRY_CHOL)) return -1;
    const double ss = M.bs + M.c0s, stt = M.bt + M.c0t, sm = M.bm + M.c0m;   // marginal sills
    double sill, nug;
    switch (km.st_model) {
    case STIF_ST_SUM: sill = ss + stt; nug = M.bs + M.bt; break;
    case STIF_ST_PRODUCT: sill = ss * stt; nug = M.bs * M.bt; break;
    case STIF_ST_PRODUCT_SUM: sill = M.k * (ss * stt) + ss * stt; nug = M.k * (M.bs * M.bt) + M.bs * M.bt; break;
    case STIF_ST_METRIC: sill = sm; nug = M.bm; break;
    default: sill = (ss + stt) + sm; nug = (M.bs + M.bt) + M.bm; break;
    }
    if (!(isfinite(sill) && isfinite(nug) && sill - nug > 0.0)) return -1;
    X.csill = sill;
    X.cdiag = sill - nug;
    X.tol = 1e-10 * X.cdiag;
    // polynomial fast path: all-spherical and no cancelling nuggets (with cancellation the
    // verbatim formula reproduces the reference's rounding, the regrouped one would not)
    const bool sph = valid && (!use_s || M.ks == STIF_M_SPHERICAL) &&
                     (!use_t || M.kt == STIF_M_SPHERICAL) && (!use_m || M.km == STIF_M_SPHERICAL);
    if (!sph) {
        X.mode = 1;
        return CHOL_GENERIC;
    }
    X.is = M.is; X.it = M.it; X.im = M.im;
    X.ncs = -M.c0s; X.nct = -M.c0t; X.ncm = -M.c0m;
    X.ani2 = M.ani2;
    X.cb = sill - nug;
    // neighbour pairs are at most 2 smax apart in space and tmax in time (both inside the
    // target's causal window); if every marginal's range covers that, q(u) needs no clamp
    {
        const double hmax = 2.0 * smax, dmax = sqrt(4.0 * smax * smax + M.ani2 * tmax * tmax);
        bool noclamp = true;
        if (use_s) noclamp = noclamp && hmax * (1.0 + 1e-9) <= M.rs;
        if (use_t) noclamp = noclamp && tmax * (1.0 + 1e-9) <= M.rt;
        if (use_m) noclamp = noclamp && dmax * (1.0 + 1e-9) <= M.rm;
        X.clamp = noclamp ? 0 : 1;     // NaN / inf radii compare false -> clamp
        // 0: polynomial, every pair inside every range; 1: regrouped with range clamps;
        // 2: the reference's expression per marginal (cancelling nuggets) with the fast sqrt
        X.mode = plain ? X.clamp : 2;
    }
    {
        auto pa = [](double c0, double inv_r) { return -c0 * 1.5 * inv_r; };
        auto pb = [](double c0, double inv_r) { return -c0 * 0.5 * inv_r * inv_r * inv_r; };
        X.pas = pa(M.c0s, M.is); X.pbs = pb(M.c0s, M.is);
        X.pat = pa(M.c0t, M.it); X.pbt = pb(M.c0t, M.it);
        X.pam = pa(M.c0m, M.im); X.pbm = pb(M.c0m, M.im);
    }
    return km.st_model == STIF_ST_SUM ? CHOL_SUM_SPH
         : km.st_model == STIF_ST_METRIC ? CHOL_METRIC_SPH : CHOL_SUM_METRIC_SPH;
}

}  // namespace krig
}  // namespace stif

using namespace stif;
using namespace stif::krig;

namespace {

enum Scratch {
    S_GRIDHDR = 0, S_OKEYS, S_OIDX, S_OKEYS_S, S_OIDX_S, S_CUB, S_OBS_SORTED, S_CELL_START,
    S_TKEYS, S_TFLAGS, S_SEL_IDX, S_SEL_GAMMA, S_STAGE, S_EVAL, S_FAIL, S_MINNORM
};

int check_ctx(stif_ctx *c, const char *who)
{
    if (!c) {
        set_error("%s: ctx is NULL", who);
        return STIF_EINVAL;
    }
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) {
        set_error("%s: cudaSetDevice(%d): %s", who, c->device, cudaGetErrorString(e));
        return STIF_ECUDA;
    }
    return STIF_OK;
}

// The call's statistics travel to a pinned host buffer behind the results (no extra round trip).
int stats_to_host(stif_ctx *c, cudaStream_t st)
{
    if (!c->k_stats_pinned)
        STIF_CUDA(cudaHostAlloc(&c->k_stats_pinned, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    STIF_CUDA(cudaMemcpyAsync(c->k_stats_pinned, c->k_stats.p, 8 * sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, st));
    return STIF_OK;
}

// The reference raises np.linalg.LinAlgError when a kriging system holds NaN / Inf (numba's
// lstsq checks the matrix, stif/predictor.py:587).  The outputs are complete (NaN for those
// targets); the code tells the caller.
int nonfinite_check(stif_ctx *c)
{
    const unsigned long long bad = c->k_stats_pinned ? c->k_stats_pinned[5] : 0ull;
    if (bad) {
        set_error("stif_kriging_predict: %llu kriging system(s) contain NaN or Inf "
                  "(check the coordinates, the times and the variogram-model parameters)", bad);
        return STIF_ENONFINITE;
    }
    return STIF_OK;
}

int64_t max_cells_for(int64_t n_obs)
{
    int64_t mc = 2 * n_obs;
    if (mc < 4096) mc = 4096;
    if (mc > (1 << 24)) mc = (1 << 24);
    return mc;
}

// Build (or reuse) the space-time grid for the current observation set and radii.
int ensure_grid(stif_ctx *c, double smax, double tmax, cudaStream_t st)
{
    const int64_t n = c->n_obs;
    if (c->grid_nobs == n && c->grid_smax == smax && c->grid_tmax == tmax) return STIF_OK;
    const int64_t max_cells = max_cells_for(n);
    const double *ox = c->k_obs.as<double>();
    const double *oy = ox + n, *ot = oy + n;
    STIF_TRY(c->k_scratch[S_GRIDHDR].reserve(6 * 8 + sizeof(GridDims)));
    STIF_TRY(c->k_scratch[S_OKEYS].reserve(4 * (size_t)n));
    STIF_TRY(c->k_scratch[S_OIDX].reserve(4 * (size_t)n));
    STIF_TRY(c->k_scratch[S_OKEYS_S].reserve(4 * (size_t)n));
    STIF_TRY(c->k_scratch[S_OIDX_S].reserve(4 * (size_t)n));
    STIF_TRY(c->k_scratch[S_OBS_SORTED].reserve(sizeof(double4) * (size_t)n));
    STIF_TRY(c->k_scratch[S_CELL_START].reserve(4 * (size_t)(max_cells + 2)));
    double *bbox = c->k_scratch[S_GRIDHDR].as<double>();
    GridDims *gd = reinterpret_cast<GridDims *>(bbox + 6);
    // min keys start at all-ones, max keys at zero
    STIF_CUDA(cudaMemsetAsync(bbox, 0xff, 3 * 8, st));
    STIF_CUDA(cudaMemsetAsync(bbox + 3, 0x00, 3 * 8, st));
    const int tb = 256;
    int nb = (int)((n + tb - 1) / tb);
    if (nb > c->sm_count * 8) nb = c->sm_count * 8;
    if (nb < 1) nb = 1;
    bbox_kernel<<<nb, tb, 0, st>>>(ox, oy, ot, n, bbox);
    double sub[3] = {GRID_SUB_X, GRID_SUB, GRID_SUB};
    if (const char *e = getenv("STIF_GRID_SUB")) sscanf(e, "%lf,%lf,%lf", &sub[0], &sub[1], &sub[2]);
    grid_dims_kernel<<<1, 1, 0, st>>>(bbox, smax, tmax, max_cells, sub[0], sub[1], sub[2], gd);
    const unsigned nbn = (unsigned)((n + tb - 1) / tb);
    unsigned *keys = c->k_scratch[S_OKEYS].as<unsigned>(), *idx = c->k_scratch[S_OIDX].as<unsigned>();
    unsigned *keys_s = c->k_scratch[S_OKEYS_S].as<unsigned>(), *idx_s = c->k_scratch[S_OIDX_S].as<unsigned>();
    obs_keys_kernel<<<nbn, tb, 0, st>>>(ox, oy, ot, n, gd, keys, idx);
    size_t tmp = 0;
    STIF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys, keys_s, idx, idx_s, (int)n, 0, 25, st));
    STIF_TRY(c->k_scratch[S_CUB].reserve(tmp));
    STIF_CUDA(cub::DeviceRadixSort::SortPairs(c->k_scratch[S_CUB].p, tmp, keys, keys_s, idx, idx_s,
                                              (int)n, 0, 25, st));
    obs_gather_kernel<<<nbn, tb, 0, st>>>(ox, oy, ot, idx_s, n, c->k_scratch[S_OBS_SORTED].as<double4>());
    cell_start_kernel<<<(unsigned)((max_cells + 1 + tb - 1) / tb), tb, 0, st>>>(
        keys_s, n, gd, max_cells, c->k_scratch[S_CELL_START].as<unsigned>());
    STIF_CUDA(cudaGetLastError());
    c->grid_nobs = n;
    c->grid_smax = smax;
    c->grid_tmax = tmax;
    return STIF_OK;
}

}  // namespace

extern "C" int stif_kriging_set_obs(stif_ctx *c, const double *x, const double *y, const double *t,
                                    const double *resid, int64_t n, int on_device)
{
    STIF_TRY(check_ctx(c, "stif_kriging_set_obs"));
    if (n < 0 || n >= 0x7fffffff || (n > 0 && (!x || !y || !t || !resid))) {
        set_error("stif_kriging_set_obs: bad arguments (n=%lld)", (long long)n);
        return STIF_EINVAL;
    }
    STIF_TRY(c->k_obs.reserve(sizeof(double) * 4 * (size_t)(n > 0 ? n : 1)));
    double *d = c->k_obs.as<double>();
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t bytes = sizeof(double) * (size_t)n;
    if (n > 0) {
        STIF_CUDA(cudaMemcpyAsync(d, x, bytes, kind, c->stream));
        STIF_CUDA(cudaMemcpyAsync(d + n, y, bytes, kind, c->stream));
        STIF_CUDA(cudaMemcpyAsync(d + 2 * n, t, bytes, kind, c->stream));
        STIF_CUDA(cudaMemcpyAsync(d + 3 * n, resid, bytes, kind, c->stream));
    }
    STIF_TRY(c->k_obs4.reserve(sizeof(double4) * (size_t)(n > 0 ? n : 1)));
    if (n > 0) {
        obs_interleave_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
            d, d + n, d + 2 * n, d + 3 * n, n, c->k_obs4.as<double4>());
        STIF_CUDA(cudaGetLastError());
    }
    STIF_CUDA(cudaStreamSynchronize(c->stream));
    c->n_obs = n;
    c->grid_nobs = -1;  // invalidate the grid
    return STIF_OK;
}

extern "C" int stif_kriging_set_residuals(stif_ctx *c, const double *resid, int64_t n, int on_device)
{
    STIF_TRY(check_ctx(c, "stif_kriging_set_residuals"));
    if (c->n_obs <= 0 || n != c->n_obs || !resid) {
        set_error("stif_kriging_set_residuals: %lld residuals for %lld bound observations",
                  (long long)n, (long long)c->n_obs);
        return n != c->n_obs ? STIF_ESTATE : STIF_EINVAL;
    }
    double *d = c->k_obs.as<double>() + 3 * n;
    STIF_CUDA(cudaMemcpyAsync(d, resid, sizeof(double) * (size_t)n,
                              on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    obs_resid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d, n, c->k_obs4.as<double4>());
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaStreamSynchronize(c->stream));
    return STIF_OK;
}

extern "C" int stif_kriging_set_model(stif_ctx *c, int st_model, int space_model, int time_model,
                                      int metric_model, const double *params, int nparams)
{
    STIF_TRY(check_ctx(c, "stif_kriging_set_model"));
    static const int need[5] = {6, 6, 7, 4, 10};
    if (st_model < 0 || st_model > 4) {
        set_error("stif_kriging_set_model: unknown space-time model id %d", st_model);
        return STIF_EINVAL;
    }
    const int models[3] = {space_model, time_model, metric_model};
    for (int m : models) {
        if (m != STIF_M_SPHERICAL && m != STIF_M_GAUSSIAN) {
            set_error("stif_kriging_set_model: unknown marginal model id %d", m);
            return STIF_EINVAL;
        }
    }
    if (!params || nparams != need[st_model]) {
        set_error("stif_kriging_set_model: model %d needs %d parameters, got %d", st_model,
                  need[st_model], nparams);
        return STIF_EINVAL;
    }
    c->model.st_model = st_model;
    c->model.space_model = space_model;
    c->model.time_model = time_model;
    c->model.metric_model = metric_model;
    c->model.nparams = nparams;
    for (int i = 0; i < nparams; ++i) c->model.x[i] = params[i];
    return STIF_OK;
}

extern "C" int stif_variogram_model_eval(stif_ctx *c, const double *h, const double *t, int64_t n,
                                         double *out)
{
    STIF_TRY(check_ctx(c, "stif_variogram_model_eval"));
    if (c->model.st_model < 0) {
        set_error("stif_variogram_model_eval: no variogram model set (fit a variogram model first)");
        return STIF_ESTATE;
    }
    if (n < 0 || (n > 0 && (!h || !t || !out))) {
        set_error("stif_variogram_model_eval: bad arguments");
        return STIF_EINVAL;
    }
    if (n == 0) return STIF_OK;
    Model M;
    STIF_TRY(build_model(c->model, M));
    STIF_TRY(c->k_scratch[S_EVAL].reserve(3 * sizeof(double) * (size_t)n));
    double *dh = c->k_scratch[S_EVAL].as<double>(), *dt = dh + n, *dout = dt + n;
    STIF_CUDA(cudaMemcpyAsync(dh, h, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    STIF_CUDA(cudaMemcpyAsync(dt, t, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    model_eval_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(M, dh, dt, n, dout);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaMemcpyAsync(out, dout, 8 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    STIF_CUDA(cudaStreamSynchronize(c->stream));
    return STIF_OK;
}

extern "C" int stif_kriging_predict_dev(stif_ctx *c, const double *tx, const double *ty,
                                        const double *tt, const int64_t *leave_out, int64_t n_targets,
                                        int min_pts, int max_pts, double smax, double tmax, int flags,
                                        double *mean_out, double *std_out, float *w_out,
                                        float *gamma_out, uint32_t *idx_out, int32_t *n_used_out,
                                        void *stream_v)
{
    STIF_TRY(check_ctx(c, "stif_kriging_predict"));
    if (c->model.st_model < 0) {
        set_error("No Kriging function defined. Did you fit a variogram model?");  // predictor.py:852-854
        return STIF_ESTATE;
    }
    if (c->n_obs <= 0) {
        set_error("stif_kriging_predict: no observations set (stif_kriging_set_obs)");
        return STIF_ESTATE;
    }
    if (max_pts < 1 || max_pts > KMAX) {
        set_error("stif_kriging_predict: max_kriging_points must be in [1, %d], got %d", KMAX, max_pts);
        return STIF_EINVAL;
    }
    if (n_targets < 0 || (n_targets > 0 && (!tx || !ty || !tt || !mean_out || !std_out))) {
        set_error("stif_kriging_predict: bad arguments");
        return STIF_EINVAL;
    }
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : c->stream;
    c->last_stream = st;
    cudaEvent_t *ev = c->ev_override ? c->ev_override : c->ev;
    if (!c->ev_override && !c->chunk_ev.empty()) {     // phases of an earlier pipelined call
        for (auto e : c->chunk_ev) cudaEventDestroy(e);
        c->chunk_ev.clear();
    }
    STIF_TRY(c->k_stats.reserve(8 * sizeof(unsigned long long)));
    if (!c->keep_stats) STIF_CUDA(cudaMemsetAsync(c->k_stats.p, 0, 8 * sizeof(unsigned long long), st));
    STIF_CUDA(cudaEventRecord(ev[0], st));
    c->n_phase = 3;
    if (n_targets == 0) {
        for (int i = 1; i <= 3; ++i) STIF_CUDA(cudaEventRecord(ev[i], st));
        return STIF_OK;
    }
    Model M;
    STIF_TRY(build_model(c->model, M));
    STIF_TRY(ensure_grid(c, smax, tmax, st));
    STIF_CUDA(cudaEventRecord(ev[1], st));

    const int k = max_pts;
    const int nr_pad = (k + 1 + 7) / 8 * 8;      // rows of the padded system
    const int nc_pad = (k + 2 + 7) / 8 * 8;      // columns incl. the right-hand side
    const int ld = nc_pad + 1;                   // odd leading dimension
    const int cta_doubles = nr_pad * ld + 16 + 5 * (k + 1) + 1 + 4 + (k + 2) / 2 + 1;
    const size_t solve_smem = sizeof(double) * (size_t)cta_doubles;

    CholParams CX;
    const int chol_var = chol_plan(c->model, M, flags, smax, tmax, CX);
    CX.obs4 = c->k_obs4.as<double4>();
    // pivoted-LU assembly fast path: product / product_sum, spherical marginals, every neighbour
    // pair inside both ranges (same bound as chol_plan)
    int lu_var = LU_GENERIC;
    if ((M.st == STIF_ST_PRODUCT || M.st == STIF_ST_PRODUCT_SUM) && M.ks == STIF_M_SPHERICAL &&
        M.kt == STIF_M_SPHERICAL && M.rs > 0.0 && M.rt > 0.0 && 2.0 * smax * (1.0 + 1e-9) <= M.rs &&
        tmax * (1.0 + 1e-9) <= M.rt && isfinite(M.is) && isfinite(M.it))
        lu_var = (M.st == STIF_ST_PRODUCT_SUM) ? LU_PRODUCT_SUM_SPH : LU_PRODUCT_SPH;
    if (!c->k_pairtab.p) {
        std::vector<unsigned> tab;
        build_pair_table(tab);
        STIF_TRY(c->k_pairtab.reserve(tab.size() * sizeof(unsigned)));
        STIF_CUDA(cudaMemcpy(c->k_pairtab.p, tab.data(), tab.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    }
    CX.pairtab = c->k_pairtab.as<unsigned>();
    // symmetric indefinite systems (models that are not valid variograms): Bunch-Kaufman LDL'
    // (krig_ldl.cuh); STIF_KRIG_TRY_LDL sends every model through it, STIF_KRIG_FORCE_LU none
    const bool use_ldl = !(flags & STIF_KRIG_FORCE_LU) && (chol_var < 0 || (flags & STIF_KRIG_TRY_LDL));
    const bool use_chol = chol_var >= 0 && !use_ldl;
    c->k_chol_enabled = (use_chol ? 1 : 0) | (use_ldl ? 2 : 0);
    const int ldl_kp = (k + 1) & ~1;
    int ldl_adbl = (k + 2) * (k + 3) / 2;
    if (ldl_adbl < 5 * ldl_kp) ldl_adbl = 5 * ldl_kp;      // the epilogue's arrays alias the matrix
    ldl_adbl = (ldl_adbl + 1) & ~1;
    const size_t ldl_smem = sizeof(double) * ((size_t)ldl_adbl + 3 * (size_t)ldl_kp + 24);
    const int chol_nt = (k + 2 + 7) / 8, chol_kp = (k + 1) & ~1;
    const size_t chol_smem = sizeof(double) * ((size_t)chol_nt * (chol_nt + 1) / 2 * 64 + 5 * (size_t)chol_kp) +
                             sizeof(unsigned) * (size_t)chol_kp;

    // chunk so that the selection scratch stays below ~16 GB
    const size_t per_target = (size_t)k * 12 + 32;
    int64_t chunk = (int64_t)((16ull << 30) / per_target);
    if (chunk > n_targets) chunk = n_targets;
    if (chunk > 0x7fffff00) chunk = 0x7fffff00;
    STIF_TRY(c->k_scratch[S_TKEYS].reserve(4 * 4 * (size_t)chunk));
    STIF_TRY(c->k_scratch[S_TFLAGS].reserve(4 * 4 * (size_t)chunk));
    STIF_TRY(c->k_scratch[S_SEL_IDX].reserve(4 * (size_t)chunk * k));
    STIF_TRY(c->k_scratch[S_SEL_GAMMA].reserve(8 * (size_t)chunk * k));
    STIF_TRY(c->k_scratch[S_FAIL].reserve(4 * (2 * (size_t)chunk + 2)));
    unsigned *tkeys = c->k_scratch[S_TKEYS].as<unsigned>();
    unsigned *tidx = tkeys + chunk, *tkeys_s = tidx + chunk, *tidx_s = tkeys_s + chunk;
    int *sel_n = c->k_scratch[S_TFLAGS].as<int>();
    unsigned *cand_n = reinterpret_cast<unsigned *>(sel_n + chunk);
    unsigned *exam_n = cand_n + chunk, *singular = exam_n + chunk;
    double *bbox = c->k_scratch[S_GRIDHDR].as<double>();
    GridDims *gd = reinterpret_cast<GridDims *>(bbox + 6);
    const double *ox = c->k_obs.as<double>();
    const int64_t n_obs = c->n_obs;

    for (int64_t c0 = 0; c0 < n_targets; c0 += chunk) {
        const int nc = (int)((n_targets - c0 < chunk) ? (n_targets - c0) : chunk);
        const int tb = 256;
        target_keys_kernel<<<(nc + tb - 1) / tb, tb, 0, st>>>(tx + c0, ty + c0, tt + c0, nc, gd, tkeys, tidx);
        size_t tmp = 0;
        STIF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, tkeys, tkeys_s, tidx, tidx_s, nc, 0, 25, st));
        STIF_TRY(c->k_scratch[S_CUB].reserve(tmp));
        STIF_CUDA(cub::DeviceRadixSort::SortPairs(c->k_scratch[S_CUB].p, tmp, tkeys, tkeys_s, tidx,
                                                  tidx_s, nc, 0, 25, st));
        STIF_CUDA(cudaMemsetAsync(singular, 0, 4 * (size_t)nc, st));

        SearchParams SP;
        SP.obs = c->k_scratch[S_OBS_SORTED].as<double4>();
        SP.cell_start = c->k_scratch[S_CELL_START].as<unsigned>();
        SP.gd = gd;
        SP.tx = tx;
        SP.ty = ty;
        SP.tt = tt;
        SP.leave_out = reinterpret_cast<const long long *>(leave_out);
        SP.order = tidx_s;
        SP.c0 = c0;
        SP.n = nc;
        SP.smax2 = smax * smax;   // space_dist_max**2, data.py:207
        SP.tmax = tmax;
        SP.k = k;
        SP.min_pts = min_pts;
        SP.M = M;
        SP.sel_n = sel_n;
        SP.sel_idx = c->k_scratch[S_SEL_IDX].as<unsigned>();
        SP.sel_gamma = c->k_scratch[S_SEL_GAMMA].as<double>();
        SP.cand_n = cand_n;
        SP.exam_n = exam_n;
        {
            const bool sph3 = M.ks == STIF_M_SPHERICAL && M.kt == STIF_M_SPHERICAL && M.km == STIF_M_SPHERICAL;
            const unsigned sgrid = (nc + SEARCH_WARPS - 1) / SEARCH_WARPS;
            if (sph3 && M.st == STIF_ST_SUM_METRIC) search_kernel<1><<<sgrid, SEARCH_WARPS * 32, 0, st>>>(SP);
            else if (sph3 && M.st == STIF_ST_PRODUCT_SUM) search_kernel<2><<<sgrid, SEARCH_WARPS * 32, 0, st>>>(SP);
            else search_kernel<0><<<sgrid, SEARCH_WARPS * 32, 0, st>>>(SP);
        }
        STIF_CUDA(cudaGetLastError());
        if (c0 + chunk >= n_targets) STIF_CUDA(cudaEventRecord(ev[2], st));

        SolveParams QP;
        QP.ox = ox;
        QP.oy = ox + n_obs;
        QP.ot = ox + 2 * n_obs;
        QP.oresid = ox + 3 * n_obs;
        QP.order = tidx_s;
        QP.c0 = c0;
        QP.n = nc;
        QP.k = k;
        QP.ld = ld;
        QP.flags = flags;
        QP.M = M;
        QP.sel_n = sel_n;
        QP.sel_idx = SP.sel_idx;
        QP.sel_gamma = SP.sel_gamma;
        QP.mean = mean_out;
        QP.std = std_out;
        QP.w_out = w_out;
        QP.g_out = gamma_out;
        QP.idx_out = idx_out;
        QP.nused_out = n_used_out;
        QP.singular = singular;
        QP.count = nullptr;
        QP.obs4 = c->k_obs4.as<double4>();
        QP.pairtab = c->k_pairtab.as<unsigned>();
        QP.pbs = M.bs; QP.pas = 1.5 * M.c0s * M.is; QP.pcs = 0.5 * M.c0s * M.is * M.is * M.is;
        QP.pbt = M.bt; QP.pat = 1.5 * M.c0t * M.it; QP.pct = 0.5 * M.c0t * M.it * M.it * M.it;
        // ---- covariance-form Cholesky (valid models), then the pivoted LU for the rest ----
        unsigned *fail_list = c->k_scratch[S_FAIL].as<unsigned>();
        unsigned *fail_count = fail_list + chunk;
        QP.sing_list = fail_count + 1;
        QP.sing_count = QP.sing_list + chunk;
        STIF_CUDA(cudaMemsetAsync(QP.sing_count, 0, sizeof(unsigned), st));
        if (use_ldl) {
            STIF_CUDA(cudaMemsetAsync(fail_count, 0, sizeof(unsigned), st));
            CX.fail_list = fail_list;
            CX.fail_count = fail_count;
            QP.count = nullptr;
#define STIF_LDL_V(NG_, W_, LUV_)                                                                  \
    do {                                                                                           \
        auto kern = solve_ldl_kernel<NG_, W_, LUV_>;                                               \
        STIF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                       (int)ldl_smem));                                            \
        int occ = 0;                                                                               \
        STIF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W_, ldl_smem));   \
        if (occ < 1) occ = 1;                                                                      \
        int64_t gsz = (int64_t)occ * c->sm_count;                                                  \
        if (gsz > nc) gsz = nc;                                                                    \
        kern<<<(unsigned)gsz, 32 * W_, ldl_smem, st>>>(QP, CX, ldl_adbl);                          \
    } while (0)
#define STIF_LDL(NG_, W_)                                                                          \
    do {                                                                                           \
        if (lu_var == LU_PRODUCT_SUM_SPH) STIF_LDL_V(NG_, W_, LU_PRODUCT_SUM_SPH);                 \
        else if (lu_var == LU_PRODUCT_SPH) STIF_LDL_V(NG_, W_, LU_PRODUCT_SPH);                    \
        else STIF_LDL_V(NG_, W_, LU_GENERIC);                                                      \
    } while (0)
            // k + 2 rows in groups of 32; one warp per target while many targets fit an SM, one warp
            // per row group when the matrix (42 KB at k = 100) leaves room for five targets only
            {
                const char *ew = getenv("STIF_LDL_WARPS");     // experiments: warps per target
                const int wsel = ew ? atoi(ew) : 0;
                if (k <= 30) STIF_LDL(1, 1);
                else if (k <= 62) { if (wsel == 2) STIF_LDL(2, 2); else STIF_LDL(2, 1); }
                else if (k <= 94) STIF_LDL(3, 3);
                else if (k <= 126) { if (wsel == 2) STIF_LDL(4, 2); else if (wsel == 1) STIF_LDL(4, 1); else STIF_LDL(4, 4); }
                else STIF_LDL(5, 5);
            }
#undef STIF_LDL
#undef STIF_LDL_V
            STIF_CUDA(cudaGetLastError());
            QP.order = fail_list;
            QP.count = fail_count;
        } else if (use_chol) {
            STIF_CUDA(cudaMemsetAsync(fail_count, 0, sizeof(unsigned), st));
            CX.fail_list = fail_list;
            CX.fail_count = fail_count;
            QP.count = nullptr;
#define STIF_CHOL(VAR_, ZR_, MODE_)                                                                \
    do {                                                                                           \
        auto kern = solve_chol_kernel<VAR_, ZR_, MODE_>;                                           \
        STIF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                       (int)chol_smem));                                           \
        int occ = 0;                                                                               \
        STIF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, chol_smem));       \
        if (occ < 1) occ = 1;                                                                      \
        int64_t gsz = (int64_t)occ * c->sm_count;                                                  \
        if (gsz > nc) gsz = nc;                                                                    \
        kern<<<(unsigned)gsz, 32, chol_smem, st>>>(QP, CX);                                        \
    } while (0)
#define STIF_CHOL_M(VAR_, ZR_)                                                                     \
    do {                                                                                           \
        if (CX.mode == 0) STIF_CHOL(VAR_, ZR_, 0);                                                 \
        else if (CX.mode == 1) STIF_CHOL(VAR_, ZR_, 1);                                            \
        else STIF_CHOL(VAR_, ZR_, 2);                                                              \
    } while (0)
#define STIF_CHOL_Z(VAR_)                                                                          \
    do {                                                                                           \
        if (k <= 64) STIF_CHOL_M(VAR_, 2);                                                         \
        else STIF_CHOL_M(VAR_, 4);                                                                 \
    } while (0)
            switch (chol_var) {
            case CHOL_SUM_SPH: STIF_CHOL_Z(CHOL_SUM_SPH); break;
            case CHOL_METRIC_SPH: STIF_CHOL_Z(CHOL_METRIC_SPH); break;
            case CHOL_SUM_METRIC_SPH: STIF_CHOL_Z(CHOL_SUM_METRIC_SPH); break;
            default:
                if (k <= 64) STIF_CHOL(CHOL_GENERIC, 2, 1);
                else STIF_CHOL(CHOL_GENERIC, 4, 1);
                break;
            }
#undef STIF_CHOL_Z
#undef STIF_CHOL_M
#undef STIF_CHOL
            STIF_CUDA(cudaGetLastError());
            QP.order = fail_list;
            QP.count = fail_count;
        }
        {
#define STIF_SOLVE_V(WPT_, LUV_)                                                                   \
    do {                                                                                           \
        auto kern = solve_blk_kernel<WPT_, LUV_>;                                                  \
        STIF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                       (int)solve_smem));                                          \
        int occ = 0;                                                                               \
        STIF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WPT_ * 32, solve_smem)); \
        if (occ < 1) occ = 1;                                                                      \
        int64_t gsz = (int64_t)occ * c->sm_count;                                                  \
        if (gsz > nc) gsz = nc;                                                                    \
        kern<<<(unsigned)gsz, WPT_ * 32, solve_smem, st>>>(QP, nr_pad, nc_pad);                    \
    } while (0)
#define STIF_SOLVE(WPT_)                                                                           \
    do {                                                                                           \
        if (lu_var == LU_PRODUCT_SUM_SPH) STIF_SOLVE_V(WPT_, LU_PRODUCT_SUM_SPH);                  \
        else if (lu_var == LU_PRODUCT_SPH) STIF_SOLVE_V(WPT_, LU_PRODUCT_SPH);                     \
        else STIF_SOLVE_V(WPT_, LU_GENERIC);                                                       \
    } while (0)
            // one thread per row of the bordered system: 32 WPT >= k + 1
            if (k <= 31) STIF_SOLVE(1);
            else if (k <= 63) STIF_SOLVE(2);
            else if (k <= 127) STIF_SOLVE(4);
            else STIF_SOLVE(8);
#undef STIF_SOLVE
#undef STIF_SOLVE_V
        }
        STIF_CUDA(cudaGetLastError());
        {   // rank-deficient systems (flagged by the LU): minimum-norm solution like dgelsd
            const int nmax = k + 1;
            const size_t mn_smem = sizeof(double) * ((size_t)nmax * nmax + 10 * (size_t)nmax) +
                                   sizeof(float) * 2 * (size_t)nmax + sizeof(unsigned) * (size_t)nmax;
            int mn_grid = c->sm_count < nc ? c->sm_count : nc;
            STIF_TRY(c->k_scratch[S_MINNORM].reserve(sizeof(double) * (size_t)mn_grid * nmax * nmax));
            STIF_CUDA(cudaFuncSetAttribute(solve_minnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)mn_smem));
            solve_minnorm_kernel<<<mn_grid, MN_THREADS, mn_smem, st>>>(QP, c->k_scratch[S_MINNORM].as<double>(),
                                                                       nullptr);
            STIF_CUDA(cudaGetLastError());
        }
        stats_kernel<<<c->sm_count, 256, 0, st>>>(exam_n, cand_n, sel_n, singular, nc,
                                                 (use_chol || use_ldl) ? fail_count : nullptr,
                                                 c->k_stats.as<unsigned long long>());
        STIF_CUDA(cudaGetLastError());
    }
    STIF_CUDA(cudaEventRecord(ev[3], st));
    return STIF_OK;
}

extern "C" int stif_kriging_predict_host(stif_ctx *c, const double *tx, const double *ty,
                                         const double *tt, const int64_t *leave_out, int64_t n,
                                         int min_pts, int max_pts, double smax, double tmax, int flags,
                                         double *mean_out, double *std_out, float *w_out,
                                         float *gamma_out, uint32_t *idx_out, int32_t *n_used_out)
{
    STIF_TRY(check_ctx(c, "stif_kriging_predict"));
    if (n < 0 || (n > 0 && (!tx || !ty || !tt || !mean_out || !std_out))) {
        set_error("stif_kriging_predict: bad arguments");
        return STIF_EINVAL;
    }
    if (max_pts < 1 || max_pts > KMAX) {
        set_error("stif_kriging_predict: max_kriging_points must be in [1, %d], got %d", KMAX, max_pts);
        return STIF_EINVAL;
    }
    if (n == 0) return STIF_OK;
    const size_t N = (size_t)n, k = (size_t)max_pts;
    // layout: tx ty tt leave | mean std | w g idx nused
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    const size_t o_tx = take(8 * N), o_ty = take(8 * N), o_tt = take(8 * N);
    const size_t o_lo = take(leave_out ? 8 * N : 0);
    const size_t o_mean = take(8 * N), o_std = take(8 * N);
    const size_t o_w = take(w_out ? 4 * N * k : 0), o_g = take(gamma_out ? 4 * N * k : 0);
    const size_t o_idx = take(idx_out ? 4 * N * k : 0), o_nu = take(n_used_out ? 4 * N : 0);
    STIF_TRY(c->k_scratch[S_STAGE].reserve(off));
    char *b = c->k_scratch[S_STAGE].as<char>();
    cudaStream_t st = c->stream;
    constexpr int64_t PIPE_MIN = 600000, PIPE_CHUNK = 1 << 18;
    if (n > PIPE_MIN) {
        // Chunked two-stream pipeline: the host-to-device copy of chunk i+1 and the
        // device-to-host copy of chunk i-1 run under the kernels of chunk i.
        if (!c->stream2) STIF_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        cudaStream_t cs = c->stream2;
        const int nchunk = (int)((n + PIPE_CHUNK - 1) / PIPE_CHUNK);
        for (auto e : c->chunk_ev) cudaEventDestroy(e);
        c->chunk_ev.assign((size_t)4 * nchunk, nullptr);
        for (auto &e : c->chunk_ev) STIF_CUDA(cudaEventCreate(&e));
        std::vector<cudaEvent_t> ev_h(nchunk), ev_d(nchunk);
        for (int i = 0; i < nchunk; ++i) {
            STIF_CUDA(cudaEventCreateWithFlags(&ev_h[i], cudaEventDisableTiming));
            STIF_CUDA(cudaEventCreateWithFlags(&ev_d[i], cudaEventDisableTiming));
        }
        STIF_TRY(c->k_stats.reserve(8 * sizeof(unsigned long long)));
        STIF_CUDA(cudaMemsetAsync(c->k_stats.p, 0, 8 * sizeof(unsigned long long), st));
        int rc = STIF_OK;
        auto d2h = [&](int ci) -> int {
            const size_t o = (size_t)ci * PIPE_CHUNK, len = (size_t)((n - (int64_t)o < PIPE_CHUNK) ? n - (int64_t)o : PIPE_CHUNK);
            STIF_CUDA(cudaStreamWaitEvent(cs, ev_d[ci], 0));
            STIF_CUDA(cudaMemcpyAsync(mean_out + o, b + o_mean + 8 * o, 8 * len, cudaMemcpyDeviceToHost, cs));
            STIF_CUDA(cudaMemcpyAsync(std_out + o, b + o_std + 8 * o, 8 * len, cudaMemcpyDeviceToHost, cs));
            if (w_out) STIF_CUDA(cudaMemcpyAsync(w_out + o * k, b + o_w + 4 * o * k, 4 * len * k, cudaMemcpyDeviceToHost, cs));
            if (gamma_out) STIF_CUDA(cudaMemcpyAsync(gamma_out + o * k, b + o_g + 4 * o * k, 4 * len * k, cudaMemcpyDeviceToHost, cs));
            if (idx_out) STIF_CUDA(cudaMemcpyAsync(idx_out + o * k, b + o_idx + 4 * o * k, 4 * len * k, cudaMemcpyDeviceToHost, cs));
            if (n_used_out) STIF_CUDA(cudaMemcpyAsync(n_used_out + o, b + o_nu + 4 * o, 4 * len, cudaMemcpyDeviceToHost, cs));
            return STIF_OK;
        };
        for (int ci = 0; ci < nchunk && rc == STIF_OK; ++ci) {
            const size_t o = (size_t)ci * PIPE_CHUNK, len = (size_t)((n - (int64_t)o < PIPE_CHUNK) ? n - (int64_t)o : PIPE_CHUNK);
            cudaMemcpyAsync(b + o_tx + 8 * o, tx + o, 8 * len, cudaMemcpyHostToDevice, cs);
            cudaMemcpyAsync(b + o_ty + 8 * o, ty + o, 8 * len, cudaMemcpyHostToDevice, cs);
            cudaMemcpyAsync(b + o_tt + 8 * o, tt + o, 8 * len, cudaMemcpyHostToDevice, cs);
            if (leave_out) cudaMemcpyAsync(b + o_lo + 8 * o, leave_out + o, 8 * len, cudaMemcpyHostToDevice, cs);
            cudaEventRecord(ev_h[ci], cs);
            cudaStreamWaitEvent(st, ev_h[ci], 0);
            c->ev_override = &c->chunk_ev[(size_t)4 * ci];
            c->keep_stats = true;
            rc = stif_kriging_predict_dev(
                c, (double *)(b + o_tx) + o, (double *)(b + o_ty) + o, (double *)(b + o_tt) + o,
                leave_out ? (int64_t *)(b + o_lo) + o : nullptr, (int64_t)len, min_pts, max_pts, smax, tmax, flags,
                (double *)(b + o_mean) + o, (double *)(b + o_std) + o, w_out ? (float *)(b + o_w) + o * k : nullptr,
                gamma_out ? (float *)(b + o_g) + o * k : nullptr, idx_out ? (uint32_t *)(b + o_idx) + o * k : nullptr,
                n_used_out ? (int32_t *)(b + o_nu) + o : nullptr, st);
            c->ev_override = nullptr;
            c->keep_stats = false;
            if (rc != STIF_OK) break;
            cudaEventRecord(ev_d[ci], st);
            if (ci > 0) rc = d2h(ci - 1);
        }
        if (rc == STIF_OK) rc = d2h(nchunk - 1);
        if (rc == STIF_OK) rc = stats_to_host(c, st);
        cudaError_t e1 = cudaStreamSynchronize(cs), e2 = cudaStreamSynchronize(st);
        for (int i = 0; i < nchunk; ++i) {
            cudaEventDestroy(ev_h[i]);
            cudaEventDestroy(ev_d[i]);
        }
        if (rc != STIF_OK) return rc;
        STIF_CUDA(e1);
        STIF_CUDA(e2);
        STIF_CUDA(cudaGetLastError());
        return nonfinite_check(c);
    }
    STIF_CUDA(cudaMemcpyAsync(b + o_tx, tx, 8 * N, cudaMemcpyHostToDevice, st));
    STIF_CUDA(cudaMemcpyAsync(b + o_ty, ty, 8 * N, cudaMemcpyHostToDevice, st));
    STIF_CUDA(cudaMemcpyAsync(b + o_tt, tt, 8 * N, cudaMemcpyHostToDevice, st));
    if (leave_out) STIF_CUDA(cudaMemcpyAsync(b + o_lo, leave_out, 8 * N, cudaMemcpyHostToDevice, st));
    STIF_TRY(stif_kriging_predict_dev(
        c, (double *)(b + o_tx), (double *)(b + o_ty), (double *)(b + o_tt),
        leave_out ? (int64_t *)(b + o_lo) : nullptr, n, min_pts, max_pts, smax, tmax, flags,
        (double *)(b + o_mean), (double *)(b + o_std), w_out ? (float *)(b + o_w) : nullptr,
        gamma_out ? (float *)(b + o_g) : nullptr, idx_out ? (uint32_t *)(b + o_idx) : nullptr,
        n_used_out ? (int32_t *)(b + o_nu) : nullptr, st));
    STIF_CUDA(cudaMemcpyAsync(mean_out, b + o_mean, 8 * N, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaMemcpyAsync(std_out, b + o_std, 8 * N, cudaMemcpyDeviceToHost, st));
    if (w_out) STIF_CUDA(cudaMemcpyAsync(w_out, b + o_w, 4 * N * k, cudaMemcpyDeviceToHost, st));
    if (gamma_out) STIF_CUDA(cudaMemcpyAsync(gamma_out, b + o_g, 4 * N * k, cudaMemcpyDeviceToHost, st));
    if (idx_out) STIF_CUDA(cudaMemcpyAsync(idx_out, b + o_idx, 4 * N * k, cudaMemcpyDeviceToHost, st));
    if (n_used_out) STIF_CUDA(cudaMemcpyAsync(n_used_out, b + o_nu, 4 * N, cudaMemcpyDeviceToHost, st));
    STIF_TRY(stats_to_host(c, st));
    STIF_CUDA(cudaStreamSynchronize(st));
    return nonfinite_check(c);
}

// calc_kriging_weights (stif/predictor.py:566-593) as a stand-alone call: the ordinary-kriging
// weights of ONE neighbour set, given its coordinates and gamma-vector -- the callable the reference
// installs as Predictor._kriging_weights_function.  Solved like np.linalg.lstsq (SVD of the bordered
// matrix, minimum norm if rank-deficient) by the Jacobi kernel, in float64.  Host pointers.
extern "C" int stif_kriging_weights_host(stif_ctx *c, const double *xs, const double *ys, const double *ts,
                                         const double *gamma_vec, int k, double *w_out)
{
    STIF_TRY(check_ctx(c, "stif_kriging_weights"));
    if (c->model.st_model < 0) {
        set_error("Create variogram model function first.");      // predictor.py:559-560
        return STIF_ESTATE;
    }
    if (k < 1 || k > KMAX || !xs || !ys || !ts || !gamma_vec || !w_out) {
        set_error("stif_kriging_weights: bad arguments (k = %d, limit %d)", k, KMAX);
        return STIF_EINVAL;
    }
    Model M;
    STIF_TRY(build_model(c->model, M));
    const int nmax = k + 1;
    // one staging block: obs4[k] | gamma[k] | w64[k] | mean, std | V[nmax^2] | idx[k] | sel_n, list, count, singular
    const size_t n_d = 4 * (size_t)k + k + k + 2 + (size_t)nmax * nmax;
    STIF_TRY(c->k_scratch[S_EVAL].reserve(sizeof(double) * n_d + sizeof(unsigned) * ((size_t)k + 8)));
    double *d = c->k_scratch[S_EVAL].as<double>();
    double4 *obs4 = reinterpret_cast<double4 *>(d);
    double *dg = d + 4 * (size_t)k, *dw = dg + k, *dms = dw + k, *dV = dms + 2;
    unsigned *du = reinterpret_cast<unsigned *>(dV + (size_t)nmax * nmax);
    std::vector<double> h(5 * (size_t)k);
    std::vector<unsigned> hu((size_t)k + 8, 0u);
    for (int e = 0; e < k; ++e) {
        h[4 * e] = xs[e];
        h[4 * e + 1] = ys[e];
        h[4 * e + 2] = ts[e];
        h[4 * e + 3] = 0.0;
        h[4 * (size_t)k + e] = gamma_vec[e];
        hu[e] = (unsigned)e;
    }
    hu[k] = (unsigned)k;        // sel_n[0] = k
    hu[k + 1] = 0u;             // sing_list[0] = target 0
    hu[k + 2] = 1u;             // sing_count
    cudaStream_t st = c->stream;
    STIF_CUDA(cudaMemcpyAsync(d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, st));
    STIF_CUDA(cudaMemcpyAsync(du, hu.data(), sizeof(unsigned) * hu.size(), cudaMemcpyHostToDevice, st));
    SolveParams QP;
    memset(&QP, 0, sizeof(QP));
    QP.oresid = dg;             // read (not used) by the float32-storage epilogue only
    QP.c0 = 0;
    QP.n = 1;
    QP.k = k;
    QP.flags = 0;
    QP.M = M;
    QP.sel_n = reinterpret_cast<const int *>(du + k);
    QP.sel_idx = du;
    QP.sel_gamma = dg;
    QP.mean = dms;
    QP.std = dms + 1;
    QP.singular = du + k + 3;
    QP.sing_list = du + k + 1;
    QP.sing_count = du + k + 2;
    QP.obs4 = obs4;
    const size_t mn_smem = sizeof(double) * ((size_t)nmax * nmax + 10 * (size_t)nmax) +
                           sizeof(float) * 2 * (size_t)nmax + sizeof(unsigned) * (size_t)nmax;
    STIF_CUDA(cudaFuncSetAttribute(solve_minnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mn_smem));
    solve_minnorm_kernel<<<1, MN_THREADS, mn_smem, st>>>(QP, dV, dw);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaMemcpyAsync(w_out, dw, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaStreamSynchronize(st));
    return STIF_OK;
}

extern "C" int stif_debug_numpy_row_sum(stif_ctx *c, const void *rows, int n_rows, int k, int is_f32,
                                        void *out)
{
    STIF_TRY(check_ctx(c, "stif_debug_numpy_row_sum"));
    if (!rows || !out || n_rows < 1 || k < 1 || k > KMAX) {
        set_error("stif_debug_numpy_row_sum: bad arguments (n_rows=%d, k=%d)", n_rows, k);
        return STIF_EINVAL;
    }
    const size_t es = is_f32 ? 4 : 8, in_bytes = es * (size_t)n_rows * k, out_bytes = es * (size_t)n_rows;
    STIF_TRY(c->k_scratch[S_EVAL].reserve(in_bytes + out_bytes + 256));
    char *d_in = c->k_scratch[S_EVAL].as<char>(), *d_out = d_in + (in_bytes + 255) / 256 * 256;
    STIF_CUDA(cudaMemcpyAsync(d_in, rows, in_bytes, cudaMemcpyHostToDevice, c->stream));
    const unsigned grid = (unsigned)((n_rows + 127) / 128);
    if (is_f32)
        row_sum_kernel<float><<<grid, 128, 0, c->stream>>>((const float *)d_in, n_rows, k, (float *)d_out);
    else
        row_sum_kernel<double><<<grid, 128, 0, c->stream>>>((const double *)d_in, n_rows, k, (double *)d_out);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    STIF_CUDA(cudaStreamSynchronize(c->stream));
    return STIF_OK;
}

extern "C" int stif_kriging_stats(stif_ctx *c, uint64_t out[4])
{
    STIF_TRY(check_ctx(c, "stif_kriging_stats"));
    if (!out) {
        set_error("stif_kriging_stats: NULL argument");
        return STIF_EINVAL;
    }
    for (int i = 0; i < 4; ++i) out[i] = 0;
    if (!c->k_stats.p) return STIF_OK;
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    STIF_CUDA(cudaStreamSynchronize(st));
    unsigned long long h[4];
    STIF_CUDA(cudaMemcpy(h, c->k_stats.p, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) out[i] = h[i];
    return STIF_OK;
}

extern "C" int stif_kriging_solver_stats(stif_ctx *c, uint64_t out[2])
{
    STIF_TRY(check_ctx(c, "stif_kriging_solver_stats"));
    if (!out) {
        set_error("stif_kriging_solver_stats: NULL argument");
        return STIF_EINVAL;
    }
    out[0] = 0;
    out[1] = (uint64_t)c->k_chol_enabled;
    if (!c->k_stats.p) return STIF_OK;
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    STIF_CUDA(cudaStreamSynchronize(st));
    unsigned long long h = 0;
    STIF_CUDA(cudaMemcpy(&h, c->k_stats.as<unsigned long long>() + 4, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = h;
    return STIF_OK;
}

extern "C" int stif_fit_set_grid(stif_ctx *c, const double *bins_space, int ns, const double *bins_time,
                                 int nt, const double *empirical, const double *weights)
{
    STIF_TRY(check_ctx(c, "stif_fit_set_grid"));
    if (ns < 1 || nt < 1 || (int64_t)ns * nt > (1 << 24) || !bins_space || !bins_time || !empirical || !weights) {
        set_error("stif_fit_set_grid: bad arguments (n_space=%d n_time=%d)", ns, nt);
        return STIF_EINVAL;
    }
    const size_t g = (size_t)ns * nt;
    STIF_TRY(c->k_fit.reserve(sizeof(double) * (ns + nt + 2 * g + 1)));
    double *d = c->k_fit.as<double>();
    STIF_CUDA(cudaMemcpyAsync(d, bins_space, 8 * (size_t)ns, cudaMemcpyHostToDevice, c->stream));
    STIF_CUDA(cudaMemcpyAsync(d + ns, bins_time, 8 * (size_t)nt, cudaMemcpyHostToDevice, c->stream));
    STIF_CUDA(cudaMemcpyAsync(d + ns + nt, empirical, 8 * g, cudaMemcpyHostToDevice, c->stream));
    STIF_CUDA(cudaMemcpyAsync(d + ns + nt + g, weights, 8 * g, cudaMemcpyHostToDevice, c->stream));
    STIF_CUDA(cudaStreamSynchronize(c->stream));
    // mapped pinned landing buffer: the kernel writes the result straight into host memory
    if (!c->fit_out_host) STIF_CUDA(cudaHostAlloc(&c->fit_out_host, sizeof(double), cudaHostAllocMapped));
    c->fit_ns = ns;
    c->fit_nt = nt;
    return STIF_OK;
}

extern "C" int stif_fit_wmse(stif_ctx *c, int st_model, int space_model, int time_model, int metric_model,
                             const double *params, int nparams, double *out)
{
    STIF_TRY(check_ctx(c, "stif_fit_wmse"));
    static const int need[5] = {6, 6, 7, 4, 10};
    if (c->fit_ns < 1) {
        set_error("stif_fit_wmse: no grid set (stif_fit_set_grid)");
        return STIF_ESTATE;
    }
    if (st_model < 0 || st_model > 4 || !params || !out || nparams != need[st_model]) {
        set_error("stif_fit_wmse: bad model id / parameter count");
        return STIF_EINVAL;
    }
    KrigModel km;
    km.st_model = st_model;
    km.space_model = space_model;
    km.time_model = time_model;
    km.metric_model = metric_model;
    km.nparams = nparams;
    for (int i = 0; i < nparams; ++i) km.x[i] = params[i];
    Model M;
    STIF_TRY(build_model(km, M));
    const int ns = c->fit_ns, nt = c->fit_nt;
    const size_t g = (size_t)ns * nt;
    double *d = c->k_fit.as<double>();
    double *out_dev = nullptr;
    STIF_CUDA(cudaHostGetDevicePointer(&out_dev, c->fit_out_host, 0));
    wmse_kernel<<<1, 256, 0, c->stream>>>(M, d, ns, d + ns, nt, d + ns + nt, d + ns + nt + g, out_dev);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaStreamSynchronize(c->stream));
    *out = *c->fit_out_host;
    return STIF_OK;
}

extern "C" int stif_kriging_pairs_count(stif_ctx *c, const double *tx, const double *ty, const double *tt,
                                        const int64_t *leave_out, int64_t n, double smax, double tmax,
                                        int64_t *n_pairs)
{
    STIF_TRY(check_ctx(c, "stif_kriging_pairs_count"));
    if (!n_pairs || n < 0 || n >= 0x7fffffff || (n > 0 && (!tx || !ty || !tt))) {
        set_error("stif_kriging_pairs_count: bad arguments");
        return STIF_EINVAL;
    }
    if (c->n_obs <= 0) {
        set_error("stif_kriging_pairs_count: no observations set (stif_kriging_set_obs)");
        return STIF_ESTATE;
    }
    *n_pairs = 0;
    c->n_pairs = 0;
    if (n == 0) return STIF_OK;
    cudaStream_t st = c->stream;
    const size_t N = (size_t)n;
    STIF_TRY(c->k_scratch[S_STAGE].reserve(8 * N * 4 + 1024));
    double *dtx = c->k_scratch[S_STAGE].as<double>(), *dty = dtx + N, *dtt = dty + N;
    long long *dlo = reinterpret_cast<long long *>(dtt + N);
    STIF_CUDA(cudaMemcpyAsync(dtx, tx, 8 * N, cudaMemcpyHostToDevice, st));
    STIF_CUDA(cudaMemcpyAsync(dty, ty, 8 * N, cudaMemcpyHostToDevice, st));
    STIF_CUDA(cudaMemcpyAsync(dtt, tt, 8 * N, cudaMemcpyHostToDevice, st));
    if (leave_out) STIF_CUDA(cudaMemcpyAsync(dlo, leave_out, 8 * N, cudaMemcpyHostToDevice, st));
    STIF_TRY(ensure_grid(c, smax, tmax, st));
    STIF_TRY(c->k_pairs_cnt.reserve(16 * N + 16));
    unsigned long long *counts = c->k_pairs_cnt.as<unsigned long long>(), *offsets = counts + N;
    PairsParams P;
    P.obs = c->k_scratch[S_OBS_SORTED].as<double4>();
    P.cell_start = c->k_scratch[S_CELL_START].as<unsigned>();
    P.gd = reinterpret_cast<GridDims *>(c->k_scratch[S_GRIDHDR].as<double>() + 6);
    P.tx = dtx; P.ty = dty; P.tt = dtt;
    P.leave_out = leave_out ? dlo : nullptr;
    P.n = (int)n;
    P.smax2 = smax * smax;
    P.tmax = tmax;
    P.counts = counts;
    P.offsets = offsets;
    P.keys = nullptr;
    const unsigned grid = (unsigned)((n + 3) / 4);
    pairs_kernel<false><<<grid, 128, 0, st>>>(P);
    STIF_CUDA(cudaGetLastError());
    size_t tmp = 0;
    STIF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, counts, offsets, (int)n, st));
    STIF_TRY(c->k_scratch[S_CUB].reserve(tmp));
    STIF_CUDA(cub::DeviceScan::ExclusiveSum(c->k_scratch[S_CUB].p, tmp, counts, offsets, (int)n, st));
    unsigned long long last[2];
    STIF_CUDA(cudaMemcpyAsync(&last[0], counts + (N - 1), 8, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaMemcpyAsync(&last[1], offsets + (N - 1), 8, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaStreamSynchronize(st));
    const unsigned long long total = last[0] + last[1];
    if (total > 0) {
        STIF_TRY(c->k_pairs_a.reserve(8 * (size_t)total));
        STIF_TRY(c->k_pairs_b.reserve(16 * (size_t)total));   // sorted keys, later the unpacked pairs
        P.keys = c->k_pairs_a.as<unsigned long long>();
        pairs_kernel<true><<<grid, 128, 0, st>>>(P);
        STIF_CUDA(cudaGetLastError());
        unsigned long long *sorted = c->k_pairs_b.as<unsigned long long>();
        STIF_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, P.keys, sorted, (int64_t)total, 0, 64, st));
        STIF_TRY(c->k_scratch[S_CUB].reserve(tmp));
        STIF_CUDA(cub::DeviceRadixSort::SortKeys(c->k_scratch[S_CUB].p, tmp, P.keys, sorted, (int64_t)total, 0, 64, st));
        // unpack into k_pairs_a (reused: 8 B per key -> 16 B per pair needs its own space)
        STIF_TRY(c->k_pairs_cnt.reserve(16 * N + 16));
        STIF_CUDA(cudaStreamSynchronize(st));
    }
    c->n_pairs = (int64_t)total;
    *n_pairs = (int64_t)total;
    return STIF_OK;
}

extern "C" int stif_kriging_pairs_fetch(stif_ctx *c, int64_t *pairs_out)
{
    STIF_TRY(check_ctx(c, "stif_kriging_pairs_fetch"));
    if (c->n_pairs < 0) {
        set_error("stif_kriging_pairs_fetch: call stif_kriging_pairs_count first");
        return STIF_ESTATE;
    }
    if (c->n_pairs == 0) return STIF_OK;
    if (!pairs_out) {
        set_error("stif_kriging_pairs_fetch: NULL output");
        return STIF_EINVAL;
    }
    const int64_t total = c->n_pairs;
    cudaStream_t st = c->stream;
    // sorted keys sit in the first half of k_pairs_b; unpack into k_pairs_a (16 B per pair)
    STIF_TRY(c->k_pairs_a.reserve(16 * (size_t)total));
    pairs_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        c->k_pairs_b.as<unsigned long long>(), total, c->k_pairs_a.as<long long>());
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaMemcpyAsync(pairs_out, c->k_pairs_a.p, 16 * (size_t)total, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaStreamSynchronize(st));
    return STIF_OK;
}
