// variogram.cu -- all-pairs empirical space-time variogram / covariogram on sm_100a.
//
// Replaces the numba loop of stif/utils.py:104-190 (get_variogram) and :193-279
// (get_covariogram), all-pairs branch of _pair_index_generator (stif/utils.py:87-92).
//
// Executed arithmetic (bit-exact with the reference's LLVM fast-math code, see
// oracle/oracle_variogram.c): d2 = fma(dy,dy,rn(dx*dx)); the cut-off sqrt_rn(d2) > smax
// is evaluated as d2 > T where T is the largest double with sqrt_rn(T) <= smax (exact,
// monotone; found on the host); bins are trunc(sqrt_rn(d2)*(1/ws)), trunc(|dt|*(1/wt)).
//
// Pipeline (all on one stream, no host synchronisation):
//   1. bounding box of (x, y, t, v) -> constants of the float pre-filter and of the fixed-point
//      sums (filter_setup_kernel); radix-sort of the observation indices by the top 32 bits of the
//      time key, or by (time slab, Hilbert cell) in the space-time sort mode (CUB; plumbing)
//   2. gather into sorted SUB-TILE RECORDS of TJ observations, laid out exactly as the pair kernel
//      wants them in shared memory ({x,y}[TJ] | {t,v}[TJ] | float x-cx [TJ] | y-cy | t-ct; NaN
//      padded), so that a tile is staged by ONE bulk copy (cp.async.bulk, the TMA engine) instead
//      of per-thread loads, conversions and stores
//   3. work list of (i-tile, j-sub-tile) items: time-sorted mode: per i-tile (512 obs) the last
//      j-sub-tile that can still be inside the time window; space-time mode: explicit list from
//      exact box-gap tests in space and time.  Every other tile pair is provably outside the
//      window and is never visited ("culling")
//   4. persistent pair kernel: grid = SMs x resident CTAs.  A block claims CHUNKS of consecutive
//      items (same i-tile): the i-tile is bulk-copied once per chunk, the j-sub-tiles stream
//      through a two-stage ring of shared-memory buffers (mbarrier "full" barriers armed with the
//      expected byte count; the last warp to finish a buffer issues the bulk copy of the
//      next-but-one item into it), so the warps of a block never wait for each other inside a
//      chunk.  Per item each thread keeps 2 i-points in registers AS CENTRED FLOATS and sweeps the
//      j-points broadcast from shared memory through a float pre-filter with a rigorous error band
//      (no false negatives; 5-7 FP32 instructions per pair, issued as packed f32x2).  Pairs it
//      cannot rule out are appended (ballot compaction, 8-pair mask entries) to the warp's
//      circular queue and binned in dense batches of 32 with the reference's exact FP64
//      arithmetic: cut-off tests, bins, and shared-memory histograms of counts and of the sums AS
//      62-BIT FIXED-POINT INTEGERS in three 21-bit limbs (native 32-bit shared atomics, carries
//      on wrap-around): integer addition is associative, so the sums are bit-reproducible from
//      run to run and independent of which block binned which item.
//   5. exact 128-bit integer reduction of the per-block limb histograms, ONE rounding to float64.
// From 163 840 observations steps 1-3 run in the SPACE-TIME SORT MODE (see the block comment
// above bbox_kernel): order by (time slab, Hilbert cell), bounding boxes of the sub-tiles, explicit
// item list from exact box-gap tests in space and time.  Same kernel, same histogram.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <stdlib.h>
#include <vector>

namespace stif {
namespace vario {

constexpr int VT = 256;        // threads per block
constexpr int VR = 2;          // i-points per thread (registers)
constexpr int TI = VT * VR;    // i-tile: 512 observations
#ifndef STIF_VARIO_TJ
#define STIF_VARIO_TJ 256
#endif
constexpr int TJ = STIF_VARIO_TJ;  // j-sub-tile: 256 observations
constexpr int JPI = TI / TJ;   // j-sub-tiles per i-tile
#ifndef STIF_VARIO_QW
#define STIF_VARIO_QW 256
#endif
constexpr int QW = STIF_VARIO_QW;  // entries (uint32) of a warp's circular queue
#ifndef STIF_VARIO_NB
#define STIF_VARIO_NB 8
#endif
constexpr int NB = STIF_VARIO_NB;  // pair tests per lane and vote = bits of a queue entry's mask
constexpr int JV = NB / VR;    // j-points per queue entry
constexpr int LVR = VR == 4 ? 2 : 1, LJV = (NB == 16 ? 4 : 3) - LVR;
#ifndef STIF_VARIO_MINB
#define STIF_VARIO_MINB 4
#endif
constexpr int MINB = STIF_VARIO_MINB;  // resident CTAs per SM the kernel is compiled for (64 registers)
constexpr int VG = JV > 4 ? JV : 4;  // j-points between queue checks
constexpr int MAX_BINS = 8192; // histogram limit (16 B of shared memory per bin)
constexpr unsigned FLUSH_ITEMS = 8192;  // 2^13 items x TI*TJ (<= 2^18) pairs < 2^32: u32 counts are safe
// Items a block claims at once (runs over the same i-tile stream their j-sub-tiles through the
// two-stage ring without a block barrier).  Measured on B200 (profiles/r02_notes.md): chunks of
// 8 / 16 / 32 items are 1-8 % SLOWER than single items at 300k and 1M observations -- with four
// resident CTAs per SM the item barrier of one block is covered by the warps of the others, while
// coarser claims lengthen the tail -- so the default is 1 and the ring then serves as the landing
// buffer of the bulk copies only.
#ifndef STIF_VARIO_CHUNK
#define STIF_VARIO_CHUNK 1
#endif
constexpr int CHUNK_MAX = STIF_VARIO_CHUNK;
#ifndef STIF_VARIO_NSTAGE
#define STIF_VARIO_NSTAGE (STIF_VARIO_CHUNK > 1 ? 2 : 1)
#endif
constexpr int NSTAGE = STIF_VARIO_NSTAGE;      // j-sub-tile buffers in flight

// A sub-tile record in global memory == its image in shared memory.
constexpr int REC_XY = 0;                    // double2 {x, y} [TJ]
constexpr int REC_TV = 16 * TJ;              // double2 {t, v} [TJ]
constexpr int REC_F = 32 * TJ;               // float x - cx [TJ] | y - cy [TJ] | t - ct [TJ]
constexpr int REC_BYTES = 44 * TJ;           // 11 264 B at TJ = 256
constexpr int REC_D_BYTES = 32 * TJ;         // the FP64 part (what an i-tile needs)

// Fixed-point sums: a term val is added as x = rn(val * 2^e) + bias, 0 <= x < 2^62, split into
// limbs of 21 + 21 + 20 bits that are accumulated in uint32 shared-memory counters.  A counter
// that wraps carries 2^32 = 2^11 units of the next limb; wraps of the top counter are counted in
// a per-block uint32 in global memory (a wrap there takes >= 2^12 binned pairs of one bin).
constexpr int LIMB_BITS = 21;
constexpr unsigned LIMB_MASK = (1u << LIMB_BITS) - 1u;
constexpr unsigned LIMB_CARRY = 1u << (32 - LIMB_BITS);   // 2048
constexpr unsigned POISON = 0x80000000u;     // a non-finite term reached this bin: the sum is NaN

// Constants of the float pre-filter, written by filter_setup_kernel.
struct FilterConst {
    double cx, cy, ct;  // centre of the bounding box (subtracted before the float conversion)
    float thi;          // d2f > thi  => the pair is certainly outside the space cut-off
    float tmaxf;        // |dtf| > tmaxf => certainly outside the time cut-off
    // fixed-point sums: x = rn(val * fx_scale) + fx_bias for |val| <= fx_rmax (else: poison)
    double fx_scale;    // 2^e
    double fx_inv;      // 2^-e
    double fx_rmax;     // bound on |val| of finite inputs
    long long fx_bias;  // 0 (variogram: val >= 0) or 2^60 (covariogram: val of either sign)
};

struct Params {
    const unsigned char *recs;   // sorted, padded observations as sub-tile records (REC_BYTES each)
    const int *item_prefix;  // [nI+1] exclusive prefix of items per i-tile
    const int *item_jbeg;    // [nI]  (= I*JPI)
    int nI;
    int n_bins, n_time_bins, n_space_bins;
    int mode;
    int part, nparts;
    double tcut;     // largest d2 with sqrt_rn(d2) <= space_dist_max
    double tmax;     // time_dist_max
    double inv_ws, inv_wt;
    double mean;
    double *partial_sum;              // [grid][n_bins] FP64 sums (default mode: L2 atomics)
    unsigned *partial_limb;           // [grid][4][n_bins] deterministic mode: limbs c0, c1, c2 and the
                                      // wrap count / poison flag
    unsigned long long *partial_cnt;  // [grid][n_bins]
    unsigned long long *stats;     // [0] chunks claimed beyond the first round, [2] pairs inside
                                   // the window, [3] items visited
    const double *thr;    // [n_space_bins + 2] d2 thresholds: bin b <=> thr[b] <= d2 < thr[b+1]
    int use_thr;          // 0: evaluate sqrt/mul/trunc per binned pair (degenerate widths)
    float inv_ws_f;       // float copy of inv_ws for the first guess
    float guard;          // distance from an integer below which the float guess is re-checked
    float guard_hi;       // 1 - guard
    float ns_f;           // (float)n_space_bins
    int sorted;           // observations are time-sorted (culling on)
    const FilterConst *filter;
    // space-time sort mode: explicit item list (I << 16 | J) and the sub-tiles' bounding boxes
    const unsigned *items;   // nullptr: items are the j-ranges of item_prefix / item_jbeg
    const double *bb;        // [nJ][6] {xlo, xhi, ylo, yhi, tlo, thi} of every j-sub-tile
};

// element k of the sorted order inside the record array
__device__ __forceinline__ const double2 *rec_xy(const unsigned char *recs, int64_t k)
{
    return reinterpret_cast<const double2 *>(recs + (size_t)(k / TJ) * REC_BYTES + REC_XY) + (k % TJ);
}
__device__ __forceinline__ const double2 *rec_tv(const unsigned char *recs, int64_t k)
{
    return reinterpret_cast<const double2 *>(recs + (size_t)(k / TJ) * REC_BYTES + REC_TV) + (k % TJ);
}

__global__ void pack_keys_kernel(const double *t, int64_t n, double *keys, unsigned *idx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = t[i];
        idx[i] = (unsigned)i;
    }
}

__global__ void iota_kernel(int64_t n, unsigned *idx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (unsigned)i;
}

// Order-preserving map double -> uint64 (every real maps above 0, so zero-initialised
// atomicMax cells mean "no data yet").
__device__ __forceinline__ unsigned long long ord_enc(double d)
{
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_dec(unsigned long long e)
{
    const unsigned long long u = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)u);
}

// Gather into the sorted sub-tile records (FP64 halves + the centred float copy the sweep reads).
// Padding (k >= n): NaN coordinates never pass d2 <= tcut.
__global__ void gather_kernel(const double *x, const double *y, const double *t, const double *v,
                              const unsigned *idx, int64_t n, int64_t n_pad, unsigned char *recs,
                              const FilterConst *filter)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pad) return;
    double4 p;
    if (k < n) {
        unsigned i = idx[k];
        p = make_double4(x[i], y[i], t[i], v[i]);
    } else {
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        p = make_double4(qnan, qnan, qnan, 0.0);
    }
    unsigned char *rec = recs + (size_t)(k / TJ) * REC_BYTES;
    const int e = (int)(k % TJ);
    reinterpret_cast<double2 *>(rec + REC_XY)[e] = make_double2(p.x, p.y);
    reinterpret_cast<double2 *>(rec + REC_TV)[e] = make_double2(p.z, p.w);
    float *f = reinterpret_cast<float *>(rec + REC_F);
    f[e] = __double2float_rn(__dsub_rn(p.x, filter->cx));
    f[TJ + e] = __double2float_rn(__dsub_rn(p.y, filter->cy));
    f[2 * TJ + e] = __double2float_rn(__dsub_rn(p.z, filter->ct));
}

// Constants of the float pre-filter (one thread).  Coordinates are centred on the bounding box
// and converted to float; with X = the half extent, a float coordinate difference is within
// Ex = X * 2^-22 of the exact one (two conversions at X * 2^-24 each, the float subtraction at
// 2X * 2^-24), the float distance within E = Ex + Ey of the exact distance, and its square as
// computed (FMUL + FFMA) within a relative 2^-22 of that.  Every pair the reference keeps
// (sqrt_rn(d2) <= smax) therefore has d2f <= thi; pairs above thi are dropped in the sweep,
// everything else is decided exactly, in FP64, by the drain.  Same for the time lag.
// Non-finite extents or bounds make thi / tmaxf infinite: the filter passes everything.
__global__ void filter_setup_kernel(const unsigned long long *ext, double smax, double tmax,
                                    int mode, double mean, FilterConst *F)
{
    double lo[3], hi[3], c[3], h[3];
    for (int a = 0; a < 3; ++a) {
        const bool any = ext[2 * a] != 0ull;
        hi[a] = any ? ord_dec(ext[2 * a]) : 0.0;
        lo[a] = any ? -ord_dec(ext[2 * a + 1]) : 0.0;
        c[a] = 0.5 * lo[a] + 0.5 * hi[a];
        if (!isfinite(c[a])) c[a] = 0.0;
        h[a] = fmax(fabs(hi[a] - c[a]), fabs(lo[a] - c[a])) * (1.0 + 1e-6);
    }
    const double k22 = 2.384185791015625e-07 * (1.0 + 1e-6);  // 2^-22
    const double E = (h[0] + h[1]) * k22, Et = h[2] * k22;
    const double s = fabs(smax) * (1.0 + 1e-9) + E;
    double thi = s * s * (1.0 + 1e-6);
    double tm = (fabs(tmax) + Et) * (1.0 + 1e-6);
    const float inf = __int_as_float(0x7f800000);
    F->cx = c[0];
    F->cy = c[1];
    F->ct = c[2];
    F->thi = (isfinite(thi) && thi < 1e37) ? __double2float_ru(thi) : inf;
    F->tmaxf = (isfinite(tm) && tm < 1e37) ? __double2float_ru(tm) : inf;
    // Fixed-point scale of the sums.  ext[6], ext[7] = enc(max v), enc(max -v) over the non-NaN
    // values.  Variogram: val = rn(rn(vi - vj)^2) <= rn(rn(vmax - vmin)^2) (rounding is monotone);
    // covariogram: |val| = |rn(rn(vi - mean) rn(vj - mean))| <= max |v - mean|^2.  With
    // R = f 2^p, f in [0.5, 1), the scale 2^(60 - p) puts every |rn(val * scale)| below 2^60, and
    // one term is quantised to 2^-60 of the largest possible term (a relative 1e-18).
    double R = 0.0;
    if (ext[6] != 0ull) {
        const double vhi = ord_dec(ext[6]), vlo = -ord_dec(ext[7]);
        if (mode == 0) {
            const double d = __dsub_rn(vhi, vlo);
            R = __dmul_rn(d, d);
        } else {
            const double a = fmax(fabs(__dsub_rn(vhi, mean)), fabs(__dsub_rn(vlo, mean)));
            R = __dmul_rn(a, a);
        }
    }
    if (!(R < 1.79e308)) R = 1.79e308;          // infinite values: their terms are poison
    if (!(R > 1e-290)) R = 1e-290;              // all values equal (or denormal spread)
    int ex;
    frexp(R, &ex);                              // R = f 2^ex, f in [0.5, 1)
    F->fx_scale = ldexp(1.0, 60 - ex);
    F->fx_inv = ldexp(1.0, ex - 60);
    F->fx_rmax = R;
    F->fx_bias = mode == 0 ? 0ll : (1ll << 60);
}

// The observations are radix-sorted by the TOP 32 BITS of the order-preserving time key only
// (4 CUB passes instead of 8; the sort is launch-latency bound at 1e5 keys): positions p < q
// satisfy coarse(t_p) <= coarse(t_q), and keys that share a coarse key (relative spacing 2^-20)
// stay in input order.  Everything that reasons about a tile's time range therefore uses the
// edges of the first / last element's coarse bucket, which bound every element of the tile.
// Zero is special-cased (CUB may or may not sort -0.0 with +0.0).
__device__ __forceinline__ double coarse_lo(double t)
{
    if (t == 0.0) return -2.2250738585072014e-308;
    return ord_dec(ord_enc(t) & 0xffffffff00000000ull);
}
__device__ __forceinline__ double coarse_hi(double t)
{
    if (t == 0.0) return 2.2250738585072014e-308;
    return ord_dec(ord_enc(t) | 0x00000000ffffffffull);   // NaN for the +inf bucket: never culled
}

// One thread per i-tile: last j-sub-tile that may hold a pair with rn(tj - ti) <= tmax.
// The smallest lag between i-tile I and sub-tile J (beyond the diagonal) is at least
// rn(coarse_lo(t[J*TJ]) - coarse_hi(t_last(I))), and rounding is monotone.
__global__ void tile_range_kernel(const unsigned char *recs, int64_t n, int nI, int nJ, double tmax,
                                  int cull, int *count, int *jbeg)
{
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nI) return;
    const int jb = I * JPI;
    int jend = nJ - 1;
    if (cull) {
        int64_t last = (int64_t)I * TI + TI - 1;
        if (last > n - 1) last = n - 1;
        const double ti_max = coarse_hi(rec_tv(recs, last)->x);
        // first J in (jb+JPI-1, nJ) with lag > tmax; all later ones are also outside
        int lo = jb + JPI, hi = nJ;  // search in [lo, hi)
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            double lag = __dsub_rn(coarse_lo(rec_tv(recs, (int64_t)mid * TJ)->x), ti_max);
            if (lag > tmax) hi = mid; else lo = mid + 1;
        }
        jend = lo - 1;
    }
    if (jend > nJ - 1) jend = nJ - 1;
    int c = jend - jb + 1;
    if (c < 0) c = 0;
    count[I] = c;
    jbeg[I] = jb;
}

// Single-block exclusive scan (nI is small: n/1024).
__global__ void scan_kernel(const int *count, int nI, int *prefix)
{
    __shared__ int s_part[1024];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nI; base += 1024) {
        int i = base + threadIdx.x;
        int v = (i < nI) ? count[i] : 0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            int a = (threadIdx.x >= off) ? s_part[threadIdx.x - off] : 0;
            __syncthreads();
            s_part[threadIdx.x] += a;
            __syncthreads();
        }
        int incl = s_part[threadIdx.x];
        if (i < nI) prefix[i] = s_carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) prefix[nI] = s_carry;
}


// ---------------------------------------------------------------------------------------------
// Space-time sort mode (large n).  Any order of the observations gives the same histogram (the
// pair set is unordered; the peeled pair is fixed up from the caller's arrays), so the order is
// free to serve the culling: observations are sorted by (time slab of width tmax / 2, Morton code
// of a 256 x 256 grid over the bounding box), which makes every 256-observation sub-tile compact
// in time AND space.  A tile pair is visited only if the boxes of the two tiles can hold a pair
// inside both cut-offs; the test uses the reference's own rounded arithmetic on the box edges
// and is therefore exact: rn(a - b), rn(d * d) and fma(d, d, c) are monotone, so the smallest
// possible lag of a tile pair bounds every pair's computed lag from below.
// ---------------------------------------------------------------------------------------------

// Bounding box of the raw input: ext[0..7] = max of enc(x), enc(-x), enc(y), enc(-y), enc(t),
// enc(-t), enc(v), enc(-v) over the non-NaN values (v == nullptr: coordinates only, called first
// by the space-time sort mode whose keys do not need the values).
__global__ void bbox_kernel(const double *x, const double *y, const double *t, const double *v,
                            int64_t n, unsigned long long *ext)
{
    unsigned long long e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool coords = x != nullptr;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        if (coords) {
            const double px = x[k], py = y[k], pz = t[k];
            if (px == px) { e[0] = max(e[0], ord_enc(px)); e[1] = max(e[1], ord_enc(-px)); }
            if (py == py) { e[2] = max(e[2], ord_enc(py)); e[3] = max(e[3], ord_enc(-py)); }
            if (pz == pz) { e[4] = max(e[4], ord_enc(pz)); e[5] = max(e[5], ord_enc(-pz)); }
        }
        if (v) {
            const double pv = v[k];
            if (pv == pv) { e[6] = max(e[6], ord_enc(pv)); e[7] = max(e[7], ord_enc(-pv)); }
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (c < 6 ? !coords : v == nullptr) continue;
        const unsigned hi = (unsigned)(e[c] >> 32), lo = (unsigned)e[c];
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        const unsigned long long m = ((unsigned long long)mh << 32) | ml;
        if ((threadIdx.x & 31) == 0 && m) atomicMax(&ext[c], m);
    }
}

__device__ __forceinline__ unsigned spread8(unsigned v)   // abcdefgh -> 0a0b0c0d0e0f0g0h
{
    v = (v | (v << 4)) & 0x0f0fu;
    v = (v | (v << 2)) & 0x3333u;
    v = (v | (v << 1)) & 0x5555u;
    return v;
}

// Hilbert index of cell (x, y) on a 256 x 256 grid (16 bits): consecutive indices are always
// edge-neighbours, so runs of the sorted observations have tighter boxes than along a Z curve.
__device__ __forceinline__ unsigned hilbert8(unsigned x, unsigned y)
{
    unsigned d = 0;
#pragma unroll
    for (unsigned s = 128; s > 0; s >>= 1) {
        const unsigned rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0) {           // rotate the quadrant
            if (rx == 1) {
                x = 255u - x;
                y = 255u - y;
            }
            const unsigned t = x;
            x = y;
            y = t;
        }
    }
    return d;
}

// key = time slab << 16 | Morton(cell x, cell y).  Only the ORDER depends on these keys, never a
// result, so their arithmetic needs no care beyond being defined for every input (NaN / inf).
__global__ void spacetime_keys_kernel(const double *x, const double *y, const double *t, int64_t n,
                                      const unsigned long long *ext, double slab_w, unsigned *keys,
                                      unsigned *idx)
{
    double lo[3], inv[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const bool any = ext[2 * a] != 0ull;
        const double hi = any ? ord_dec(ext[2 * a]) : 0.0;
        lo[a] = any ? -ord_dec(ext[2 * a + 1]) : 0.0;
        const double range = hi - lo[a];
        double w = range / 256.0;                        // x, y: 256 cells over the box
        if (a == 2) {
            w = slab_w;                                  // t: slabs of tmax / 2 ...
            if (!(range / w < 65535.0)) w = range / 65535.0;   // ... unless that needs > 2^16 slabs
        }
        inv[a] = (isfinite(lo[a]) && isfinite(w) && w > 0.0) ? 1.0 / w : 0.0;
        if (!isfinite(lo[a]) || !isfinite(inv[a])) { lo[a] = 0.0; inv[a] = 0.0; }
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double fx = (x[i] - lo[0]) * inv[0], fy = (y[i] - lo[1]) * inv[1], ft = (t[i] - lo[2]) * inv[2];
    // fmin / fmax drop NaN: NaN coordinates land in cell 255 / slab 65535
    const unsigned cx = (unsigned)fmin(fmax(fx == fx ? fx : 255.0, 0.0), 255.0);
    const unsigned cy = (unsigned)fmin(fmax(fy == fy ? fy : 255.0, 0.0), 255.0);
    const unsigned sl = (unsigned)fmin(fmax(ft == ft ? ft : 65535.0, 0.0), 65535.0);
#ifdef STIF_VARIO_MORTON
    keys[i] = (sl << 16) | (spread8(cy) << 1) | spread8(cx);
#else
    keys[i] = (sl << 16) | hilbert8(cx, cy);
#endif
    idx[i] = (unsigned)i;
}

// One warp per j-sub-tile: box of its non-NaN coordinates (empty: lo = +inf, hi = -inf).
__global__ void tile_bbox_kernel(const unsigned char *recs, int nJ, double *bb)
{
    const int J = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (J >= nJ) return;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int e = lane; e < TJ; e += 32) {
        const double2 pxy = *rec_xy(recs, (int64_t)J * TJ + e);
        const double pz = rec_tv(recs, (int64_t)J * TJ + e)->x;
        lo[0] = fmin(lo[0], pxy.x); hi[0] = fmax(hi[0], pxy.x);   // fmin / fmax ignore NaN
        lo[1] = fmin(lo[1], pxy.y); hi[1] = fmax(hi[1], pxy.y);
        lo[2] = fmin(lo[2], pz); hi[2] = fmax(hi[2], pz);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int off = 16; off > 0; off >>= 1) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
            hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            bb[(size_t)J * 6 + 2 * a] = lo[a];
            bb[(size_t)J * 6 + 2 * a + 1] = hi[a];
        }
    }
}

// Smallest |a - b| the reference can compute for a in [alo, ahi], b in [blo, bhi] (0 if the
// intervals overlap): rn(blo - ahi) or rn(alo - bhi), rounding is monotone.  NaN (inf - inf)
// compares false everywhere below, i.e. "cannot be culled".
__device__ __forceinline__ double box_gap(double alo, double ahi, double blo, double bhi)
{
    double g = 0.0;
    const double g1 = __dsub_rn(blo, ahi), g2 = __dsub_rn(alo, bhi);
    if (g1 > g) g = g1;
    if (g2 > g) g = g2;
    return g;
}

// Work list of the space-time mode, one warp per i-tile: sub-tile J >= I * JPI is an item unless
// every pair of (I, J) is provably outside the time window or the space cut-off.
//   d2 >= fma(gy, gy, rn(gx gx))  (utils.py:155-160 as executed),  |dt| >= gt  (utils.py:163-165)
// FILL = false: count[I]; FILL = true: items[prefix[I] + rank] = I << 16 | J, J ascending.
template <bool FILL>
__global__ void items_kernel(const double *bb, int nI, int nJ, double tcut, double tmax,
                             int *count, const int *prefix, unsigned *items)
{
    const int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (I >= nI) return;
    double ilo[3] = {INFINITY, INFINITY, INFINITY}, ihi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int q = 0; q < JPI; ++q) {
        const int J = I * JPI + q;
        if (J < nJ) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                ilo[a] = fmin(ilo[a], bb[(size_t)J * 6 + 2 * a]);
                ihi[a] = fmax(ihi[a], bb[(size_t)J * 6 + 2 * a + 1]);
            }
        }
    }
    int total = 0;
    const int base = FILL ? prefix[I] : 0;
    for (int J0 = I * JPI; J0 < nJ; J0 += 32) {
        const int J = J0 + lane;
        bool keep = false;
        if (J < nJ) {
            keep = true;
            if (J >= I * JPI + JPI) {
                const double *b = bb + (size_t)J * 6;
                const double gx = box_gap(ilo[0], ihi[0], b[0], b[1]);
                const double gy = box_gap(ilo[1], ihi[1], b[2], b[3]);
                const double gt = box_gap(ilo[2], ihi[2], b[4], b[5]);
                const double d2min = __fma_rn(gy, gy, __dmul_rn(gx, gx));
                if (d2min > tcut || gt > tmax) keep = false;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (FILL && keep) items[base + total + __popc(bal & ((1u << lane) - 1u))] = ((unsigned)I << 16) | (unsigned)J;
        total += __popc(bal);
    }
    if (!FILL && lane == 0) count[I] = total;
}

struct Smem {
    unsigned cnt_addr;       // shared-memory address of s_cnt
    double *g_sum;           // [n_bins] GLOBAL: this block's partial FP64 sums (default mode)
    unsigned *s_cnt;         // [n_bins] pair counts of this block
    unsigned *s_lim;         // [3][n_bins] limbs of the fixed-point sums (deterministic mode)
    unsigned *g_wrap;        // [n_bins] GLOBAL: wraps of the top limb counter | POISON
    double *s_thr;           // [n_space_bins + 2]
    int n_bins;
    double fx_scale, fx_rmax;
    long long fx_bias;
};

// byte offsets into dynamic shared memory
constexpr unsigned O_I = 0;                               // i-tile: FP64 parts of its JPI records
constexpr unsigned O_J = O_I + JPI * REC_D_BYTES;         // NSTAGE j-sub-tile records
constexpr unsigned O_Q = O_J + NSTAGE * REC_BYTES;        // [VT/32][QW] one circular queue per warp
constexpr unsigned O_BAR = O_Q + (VT / 32) * QW * 4;      // mbarriers, counters
constexpr unsigned O_THR = O_BAR + 64;                    // s_thr, then s_cnt, s_lim
// O_BAR + 0: full barrier of the i-tile; + 8 (1 + s): full barrier of j stage s;
// + 32 + 4 s: warps done with stage s; + 48: next chunk (long long)

// Keep a value in a register: the result of an `asm volatile` cannot be re-derived, so the compiler
// does not rematerialise it (from %tid / %cluster_ctaid / kernel parameters) inside the hot loops.
__device__ __forceinline__ unsigned pin_u32(unsigned v)
{
    unsigned r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

__device__ __forceinline__ void st_shared_u32(unsigned addr, unsigned v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ double2 lds_d2(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double lds_d(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// ---- mbarrier + bulk copy (the TMA engine; SASS: SYNCS.*, UBLKCP) ----
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// global -> shared bulk copy, completion counted in bytes on `bar` (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// One term of a bin's sum, as a 62-bit fixed-point integer in three shared-memory limb counters.
// The three adds are independent native ATOMS.ADD; a counter that wraps (the returned old value
// shows it) carries into the next one -- once per >= 2048 adds.  Integer addition is associative:
// the total does not depend on the order in which warps, blocks or GPUs bin their pairs.
__device__ __forceinline__ void add_fixed(const Smem &S, int bin, double val)
{
    if (fabs(val) <= S.fx_rmax) {          // false for NaN / Inf
        const long long x = __double2ll_rn(__dmul_rn(val, S.fx_scale)) + S.fx_bias;   // 0 <= x < 2^62
        const unsigned l0 = (unsigned)x & LIMB_MASK;
        const unsigned l1 = (unsigned)(x >> LIMB_BITS) & LIMB_MASK;
        const unsigned l2 = (unsigned)(x >> (2 * LIMB_BITS));
        unsigned *c0 = S.s_lim + bin, *c1 = c0 + S.n_bins, *c2 = c1 + S.n_bins;
        const unsigned o0 = atomicAdd(c0, l0), o1 = atomicAdd(c1, l1), o2 = atomicAdd(c2, l2);
        unsigned w1 = (o0 + l0 < o0) ? 1u : 0u, w2 = (o1 + l1 < o1) ? 1u : 0u, w3 = (o2 + l2 < o2) ? 1u : 0u;
        if (w1 | w2 | w3) {
            if (w1) {
                const unsigned o = atomicAdd(c1, LIMB_CARRY);
                w2 += (o + LIMB_CARRY < o) ? 1u : 0u;
            }
            for (; w2; --w2) {
                const unsigned o = atomicAdd(c2, LIMB_CARRY);
                w3 += (o + LIMB_CARRY < o) ? 1u : 0u;
            }
            if (w3) atomicAdd(S.g_wrap + bin, w3);
        }
    } else {
        atomicOr(S.g_wrap + bin, POISON);
    }
}

// Bin one queued pair per lane: j-point j of the stage at byte offset jbase and i-point ii of
// the i-tile, everything in FP64 exactly as the reference executes it.  The queue holds every
// pair the float pre-filter could not rule out, so the reference's two cut-off tests are applied
// here.  Returns whether the pair is inside the window (statistics).
// Histogram update: the count is a native shared-memory atomic (ATOMS.POPC.INC).  The sum goes,
//   FX = false (default): to the block's partial histogram in GLOBAL memory with a native L2 FP64
//     atomic (REDG.E.ADD.F64, fire and forget; a shared-memory f64 add is a CAS loop on sm_100a).
//     The order of the adds -- hence the last bits of the sums -- depends on the scheduling.
//   FX = true (STIF_VARIO_DETERMINISTIC): as a fixed-point integer into shared-memory limb
//     counters (add_fixed): bit-reproducible, ~25 more instructions per binned pair.
template <int MODE, bool FX>
__device__ __forceinline__ void bin_pair(const Smem &S, const Params &P, unsigned ia, unsigned ja, bool have,
                                         unsigned &npass)
{
    // ia / ja: shared-memory byte addresses of the {x,y} halves of the i-point and the j-point
    bool inside = false;
    if (have) {
        const double2 ixy = lds_d2(ia + REC_XY), itv = lds_d2(ia + REC_TV);
        const double2 jxy = lds_d2(ja + REC_XY), jtv = lds_d2(ja + REC_TV);
        const double dx = __dsub_rn(ixy.x, jxy.x);
        const double dy = __dsub_rn(ixy.y, jxy.y);
        const double d2 = __fma_rn(dy, dy, __dmul_rn(dx, dx));
        const double lag_t = fabs(__dsub_rn(itv.x, jtv.x));
        inside = (d2 <= P.tcut) && (lag_t <= P.tmax);  // utils.py:160-165
        if (inside) {
            ++npass;                                   // statistics: pairs inside the window
            // both bin indices are bounded by the cut-offs that let the pair in (lag <= max =
            // n_bins * width), so 32-bit conversions are exact.
            // trunc(sqrt_rn(d2)*inv_ws) is monotone in d2: the bin is looked up in the exact d2
            // thresholds (host-bisected with the reference's formula) from a float guess.  The
            // guess itself is already the answer unless it lies within the float error bound
            // (relative 2^-21.4: rounding of d2, sqrt.approx, inv_ws_f, the product) of an
            // integer; `guard` = 2^-19 * n_space_bins, so ~1e-4 of the pairs take the lookup
            // (guard = 2 with degenerate bin widths: every pair takes the slow path, which then
            // evaluates the reference's expression).
            float f = __double2float_rn(d2), sq;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(f));
            const float vf = __fmul_rn(sq, P.inv_ws_f), fl = floorf(vf), fr = __fsub_rn(vf, fl);
            int bs = __float2int_rz(fminf(vf, P.ns_f));
            if (!(fr > P.guard && fr < P.guard_hi && fl < P.ns_f)) {
                if (P.use_thr) {
                    const unsigned thr = (unsigned)__cvta_generic_to_shared(S.s_thr);
                    bs = bs < 0 ? 0 : bs;
                    while (d2 < lds_d(thr + 8u * bs)) --bs;
                    while (d2 >= lds_d(thr + 8u * bs + 8u)) ++bs;
                } else {
                    bs = __double2int_rz(__dmul_rn(__dsqrt_rn(d2), P.inv_ws));
                }
            }
            const int bt = __double2int_rz(__dmul_rn(lag_t, P.inv_wt));
            if ((unsigned)bs < (unsigned)P.n_space_bins && (unsigned)bt < (unsigned)P.n_time_bins) {
                double val;
                if (MODE == 0) {
                    const double a = __dsub_rn(itv.y, jtv.y);
                    val = __dmul_rn(a, a);
                } else {
                    val = __dmul_rn(__dsub_rn(itv.y, P.mean), __dsub_rn(jtv.y, P.mean));
                }
                const int bin = bs * P.n_time_bins + bt;
                if (FX) add_fixed(S, bin, val);
                else atomicAdd(&S.g_sum[bin], val);
                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(S.cnt_addr + 4u * (unsigned)bin) : "memory");
            }
        }
    }
}

// Queue entry (32 bits): bits 0-7 = m, the set of pending pairs between the owner's 2 i-points (rows
// r of the i-tile) and 4 consecutive j-points (bit 2 h + r); bits 12-16 = owner lane; bits 18-23 =
// j / 4.  The fields sit where the drain needs them: the lowest pending pair is ffs(e) - 1, clearing it
// is e & (e - 1), (e >> 8) & 0x1f0 is the owner's byte offset in a 16-byte row of the i-tile and
// (e >> 12) & 0xfc0 the byte offset of the j group -- 10 instructions per batch less than the first
// layout (j / 4 | lane << 7 | m << 12), which the drain had to take apart and scale.
// A batch of 32 entries bins the lowest pending pair of each entry with all lanes active;
// entries with pairs left are re-queued at the tail (ballot compaction),
// so every batch but an item's last is full whatever the hit density.  The queue is circular:
// a batch frees 32 slots before it appends at most 32.
static_assert(NB == 8 && VR == 2 && TJ == VT && TJ <= 256, "queue entry layout assumes 2 i-points x 4 j-points, TJ = VT <= 256");
template <int MODE, bool FX>
__device__ __forceinline__ void drain(const Smem &S, const Params &P, unsigned ibase, unsigned jaddr,
                                      unsigned wq_addr, int lane, unsigned lt,
                                      unsigned &head, unsigned &tail, unsigned &npass, bool flush)
{
    // ibase: address of row 0, slot warp_base of the i-tile; jaddr: address of the j-sub-tile
    unsigned rd = ((head + (unsigned)lane) & (QW - 1)) << 2;      // this lane's read offset in the ring
    // full batches: every lane has an entry, no predicate on the way
    while (tail - head >= 32u) {
        __syncwarp();
        unsigned e;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(wq_addr + rd) : "memory");
        rd = (rd + 128u) & (QW * 4 - 1);
        const unsigned b = (unsigned)__ffs((int)e) - 1u;  // lowest pending pair: row b & 1, j offset b >> 1
        const unsigned ia = ibase + ((e >> 8) & 0x1f0u) + ((b & 1u) * (unsigned)REC_D_BYTES);
        const unsigned ja = jaddr + ((e >> 12) & 0xfc0u) + ((b & 6u) << 3);
        bin_pair<MODE, FX>(S, P, ia, ja, true, npass);
        const unsigned e2 = e & (e - 1u);
        const bool left = (e2 & 0xffu) != 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, left);
        __syncwarp();
        if (left) st_shared_u32(wq_addr + (((tail + __popc(bal & lt)) & (QW - 1)) << 2), e2);
        head += 32u;
        tail += __popc(bal);
    }
    // the item's last, partial batches
    while (flush && tail != head) {
        __syncwarp();
        const unsigned nv = min(tail - head, 32u);
        const bool have = (unsigned)lane < nv;
        unsigned e;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(wq_addr + (((head + lane) & (QW - 1)) << 2)) : "memory");
        const unsigned b = (unsigned)__ffs((int)e) - 1u;
        const unsigned ia = ibase + ((e >> 8) & 0x1f0u) + ((b & 1u) * (unsigned)REC_D_BYTES);
        const unsigned ja = jaddr + ((e >> 12) & 0xfc0u) + ((b & 6u) << 3);
        bin_pair<MODE, FX>(S, P, ia, ja, have, npass);
        const unsigned e2 = e & (e - 1u);
        const bool left = have && (e2 & 0xffu) != 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, left);
        __syncwarp();
        if (left) st_shared_u32(wq_addr + (((tail + __popc(bal & lt)) & (QW - 1)) << 2), e2);
        head += nv;
        tail += __popc(bal);
    }
    __syncwarp();
}

// Blackwell's packed single-precision instructions (FADD2 / FMUL2 / FFMA2: two IEEE round-to-
// nearest FP32 operations per lane and issue slot).  The sweep is issue-bound, so it tests two
// j-points per instruction: the operands are {j, j+1} pairs straight out of the SoA j-tile and
// the thread's i-point duplicated into both halves.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Float pre-filter of VR i-points x VG j-points per thread; pairs it cannot rule out are queued.
//   DIAG: the sub-tile overlaps the i-tile, keep only j > i.   TCHK: the time test is needed
//   (false when the whole item is provably inside the time window).
//   s_jf: the float rows x - cx | y - cy | t - ct of the j-sub-tile's record in shared memory.
template <bool DIAG, bool TCHK>
__device__ __forceinline__ void sweep(const float *s_jf, const f32x2 (&xf)[VR], const f32x2 (&yf)[VR],
                                      const f32x2 (&tf)[VR], const int (&ioff)[VR], float thi,
                                      float tmaxf, int jb, unsigned wq_addr, unsigned lt,
                                      unsigned tag, unsigned &tail)
{
#pragma unroll
    for (int g = 0; g < VG; g += JV) {
        const int j = jb + g;  // multiple of JV
        unsigned m = 0;
#pragma unroll
        for (int h = 0; h < JV; h += 2) {
            // {j+h, j+h+1}: 8-byte aligned pairs of the SoA rows, the same for every lane
            const float2 jx = *reinterpret_cast<const float2 *>(s_jf + j + h);
            const float2 jy = *reinterpret_cast<const float2 *>(s_jf + TJ + j + h);
            const f32x2 jx2 = pack2(jx.x, jx.y), jy2 = pack2(jy.x, jy.y);
            f32x2 jt2 = 0;
            if (TCHK) {
                const float2 jt = *reinterpret_cast<const float2 *>(s_jf + 2 * TJ + j + h);
                jt2 = pack2(jt.x, jt.y);
            }
#pragma unroll
            for (int r = 0; r < VR; ++r) {
                const f32x2 dx = sub2(xf[r], jx2), dy = sub2(yf[r], jy2);
                float d2a, d2b;
                unpack2(fma2(dy, dy, mul2(dx, dx)), d2a, d2b);
                // negated comparisons: a NaN (float overflow of absurd coordinates) is kept
                // and decided exactly in the drain
                bool pa = !(d2a > thi), pb = !(d2b > thi);
                if (TCHK) {
                    float dta, dtb;
                    unpack2(sub2(tf[r], jt2), dta, dtb);
                    pa = pa && !(fabsf(dta) > tmaxf);
                    pb = pb && !(fabsf(dtb) > tmaxf);
                }
                if (DIAG) {
                    pa = pa && (j + h > ioff[r]);
                    pb = pb && (j + h + 1 > ioff[r]);
                }
                m |= pa ? (1u << (VR * h + r)) : 0u;
                m |= pb ? (1u << (VR * (h + 1) + r)) : 0u;
            }
        }
        const bool any = m != 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, any);
        if (any)
            st_shared_u32(wq_addr + (((tail + __popc(bal & lt)) & (QW - 1)) << 2),
                          tag | m | ((unsigned)j << 16));     // j is a multiple of 4: (j / 4) << 18
        tail += __popc(bal);
    }
}

template <int MODE, bool FX>
__global__ void __launch_bounds__(VT, MINB) pair_kernel(Params P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    Smem S;
    S.s_thr = reinterpret_cast<double *>(smem_raw + O_THR);
    S.s_cnt = reinterpret_cast<unsigned *>(S.s_thr + P.n_space_bins + 2);
    S.s_lim = S.s_cnt + P.n_bins;
    S.n_bins = P.n_bins;
    S.cnt_addr = pin_u32((unsigned)__cvta_generic_to_shared(S.s_cnt));
    unsigned *s_done = reinterpret_cast<unsigned *>(smem_raw + O_BAR + 32);
    int *s_next = reinterpret_cast<int *>(smem_raw + O_BAR + 48);
    const unsigned bar_i = sbase + O_BAR, bar_j = sbase + O_BAR + 8;

    const int tid = threadIdx.x, lane = tid & 31;
    for (int b = tid; b < (FX ? 4 : 1) * P.n_bins; b += VT) S.s_cnt[b] = 0u;   // counts (and the three limbs)
    for (int b = tid; b < P.n_space_bins + 2; b += VT) S.s_thr[b] = P.thr[b];
    unsigned long long *my_cnt = P.partial_cnt + (size_t)blockIdx.x * P.n_bins;
    unsigned *my_limb = P.partial_limb + (size_t)blockIdx.x * 4 * P.n_bins;
    double *my_sum = P.partial_sum + (size_t)blockIdx.x * P.n_bins;
    for (int b = tid; b < P.n_bins; b += VT) {
        my_cnt[b] = 0ull;
        if (FX) my_limb[3 * P.n_bins + b] = 0u;
        else my_sum[b] = 0.0;
    }
    S.g_wrap = my_limb + 3 * P.n_bins;
    S.g_sum = my_sum;
    const FilterConst F = *P.filter;
    S.fx_scale = F.fx_scale;
    S.fx_rmax = F.fx_rmax;
    S.fx_bias = F.fx_bias;
    if (tid == 0) {
        mbar_init(bar_i, 1);
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(bar_j + 8 * s, 1);
            s_done[s] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const unsigned wq_addr = sbase + O_Q + (unsigned)(tid >> 5) * QW * 4;
    const unsigned lt = (1u << lane) - 1u, tag = (unsigned)lane << 12;
    unsigned ph_i = 0, ph_j[NSTAGE];
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) ph_j[s] = 0;

    const int M = P.item_prefix[P.nI];  // total items (< 2^27 in the space-time mode, int in the other)
    // Items are claimed in chunks of CH consecutive list entries (the list is ordered by i-tile,
    // so a chunk is usually one run over the same i-tile).  CH is the same for every block and
    // leaves each block >= 12 chunks, so the tail stays short; with the default CHUNK_MAX = 1 it is
    // a compile-time 1 and the run / stage loops below fold away.  This partition owns the chunks q
    // with q % nparts == part; a block starts with chunk blockIdx.x of its partition and then
    // takes the next unclaimed one (stats[0] counts the claims beyond the first round): items
    // differ in cost.  Which block bins which item only changes the order of the FP64 sums (not at
    // all the counts, and nothing in the deterministic mode).
    // (32-bit indices on purpose: the kernel is compiled for 64 registers, and every 64-bit loop
    // variable here was paid for by re-deriving addresses inside the drain.)
    int CH = 1;
    if (CHUNK_MAX > 1) {
        const long long per = (long long)M / ((long long)gridDim.x * P.nparts * 12);
        CH = per < 1 ? 1 : (per > CHUNK_MAX ? CHUNK_MAX : (int)per);
    }
    const int NQ = (M + CH - 1) / CH;
    unsigned long long npass_total = 0;
    unsigned items_done = 0, items_total = 0;
    __syncthreads();
    for (int kq = blockIdx.x;;) {
        const long long q64 = (long long)kq * P.nparts + P.part;
        if (q64 >= NQ) break;
        int w = (int)q64 * CH;
        const int wend = (w + CH < M) ? w + CH : M;
        bool first_run = true;
        while (w < wend) {
            // ---- one run: items w .. w + L - 1 share the i-tile I ----
            int I, J0 = 0, L;
            if (P.items) {   // space-time mode: explicit list
                I = (int)(P.items[w] >> 16);
                L = 1;
                while (w + L < wend && (int)(P.items[w + L] >> 16) == I) ++L;
            } else {
                // item -> (I, J): largest I with prefix[I] <= w
                int lo = 0, hi = P.nI - 1;
                while (lo < hi) {
                    int mid = (lo + hi + 1) >> 1;
                    if (P.item_prefix[mid] <= w) lo = mid; else hi = mid - 1;
                }
                I = lo;
                J0 = P.item_jbeg[I] + (int)(w - P.item_prefix[I]);
                const int iend = P.item_prefix[I + 1];
                L = (iend < wend ? iend : wend) - w;
            }
            auto item_J = [&](int c) -> int {
                return P.items ? (int)(P.items[w + c] & 0xffffu) : J0 + c;
            };

            __syncthreads();  // every warp is done with the previous run's i-tile and j stages
            if (items_done + (unsigned)L > FLUSH_ITEMS) {     // u32 counts could overflow: flush
                for (int b = tid; b < P.n_bins; b += VT) {
                    my_cnt[b] += S.s_cnt[b];
                    S.s_cnt[b] = 0u;
                }
                items_done = 0;
                __syncthreads();
            }
            if (tid == 0) {
                if (first_run) *s_next = (int)gridDim.x + (int)atomicAdd(&P.stats[0], 1ull);
                // the i-tile: the FP64 parts of its JPI records; the first NSTAGE j-sub-tiles
                mbar_expect_tx(bar_i, JPI * REC_D_BYTES);
#pragma unroll
                for (int s = 0; s < JPI; ++s)
                    bulk_g2s(sbase + O_I + s * REC_D_BYTES, P.recs + (size_t)(I * JPI + s) * REC_BYTES,
                             REC_D_BYTES, bar_i);
#pragma unroll
                for (int s = 0; s < NSTAGE; ++s) {
                    if (s < L) {
                        mbar_expect_tx(bar_j + 8 * s, REC_BYTES);
                        bulk_g2s(sbase + O_J + s * REC_BYTES, P.recs + (size_t)item_J(s) * REC_BYTES,
                                 REC_BYTES, bar_j + 8 * s);
                    }
                }
            }
            first_run = false;
            // the thread's own i-points as centred floats (duplicated into both halves), straight
            // from the records' float rows while the bulk copies are in flight
            const unsigned ibase = pin_u32(sbase + O_I + (unsigned)(tid & ~31) * 16u);
            f32x2 xf[VR], yf[VR], tf[VR];
#pragma unroll
            for (int r = 0; r < VR; ++r) {
                const int ii = r * VT + tid;
                const float *gf = reinterpret_cast<const float *>(
                    P.recs + (size_t)(I * JPI + ii / TJ) * REC_BYTES + REC_F) + (ii % TJ);
                const float fx = __ldg(gf), fy = __ldg(gf + TJ), ft = __ldg(gf + 2 * TJ);
                xf[r] = pack2(fx, fx);
                yf[r] = pack2(fy, fy);
                tf[r] = pack2(ft, ft);
            }
            mbar_wait(bar_i, ph_i);
            ph_i ^= 1u;

            for (int c = 0; c < L; ++c) {
                const int st = c % NSTAGE;
                const int J = item_J(c);
                const unsigned jbase = O_J + (unsigned)st * REC_BYTES;
                int ioff[VR];
#pragma unroll
                for (int r = 0; r < VR; ++r)
                    ioff[r] = (I * TI + tid) - J * TJ + r * VT;  // (global i) - (global j of sub-tile start)
                // 0: diagonal item; 1: time test needed; 2: the whole item is inside the time window
                // (every t of a tile lies between the coarse-bucket edges of its first and last
                // element; rounding is monotone)
                int variant = 1;
                if (J < I * JPI + JPI)
                    variant = 0;
                else if (P.items) {
                    // largest lag the two boxes allow: rn(thi_J - tlo_I) and rn(thi_I - tlo_J)
                    double ilo = INFINITY, ihi = -INFINITY;
#pragma unroll
                    for (int s = 0; s < JPI; ++s) {
                        ilo = fmin(ilo, P.bb[(size_t)(I * JPI + s) * 6 + 4]);
                        ihi = fmax(ihi, P.bb[(size_t)(I * JPI + s) * 6 + 5]);
                    }
                    const double jlo = P.bb[(size_t)J * 6 + 4], jhi = P.bb[(size_t)J * 6 + 5];
                    if (__dsub_rn(jhi, ilo) <= P.tmax && __dsub_rn(ihi, jlo) <= P.tmax) variant = 2;
                }
#pragma unroll
                for (int s = 0; s < NSTAGE; ++s)
                    if (s == st) {
                        mbar_wait(bar_j + 8 * s, ph_j[s]);
                        ph_j[s] ^= 1u;
                    }
                if (variant == 1 && !P.items && P.sorted) {
                    const double jt0 = lds_d(sbase + jbase + REC_TV), jt1 = lds_d(sbase + jbase + REC_TV + 16u * (TJ - 1));
                    const double it0 = lds_d(sbase + O_I + REC_TV);
                    const double it1 = lds_d(sbase + O_I + (JPI - 1) * REC_D_BYTES + REC_TV + 16u * (TJ - 1));
                    if (__dsub_rn(coarse_hi(jt1), coarse_lo(it0)) <= P.tmax &&
                        __dsub_rn(coarse_hi(it1), coarse_lo(jt0)) <= P.tmax)
                        variant = 2;
                }
                const float *s_jf = reinterpret_cast<const float *>(smem_raw + jbase + REC_F);

                unsigned head = 0, tail = 0, npass = 0;
                for (int jb = 0; jb < TJ; jb += VG) {
                    if (tail - head > (unsigned)(QW - (VG / JV) * 32))
                        drain<MODE, FX>(S, P, ibase, sbase + jbase, wq_addr, lane, lt, head, tail, npass, false);
                    if (variant == 2)
                        sweep<false, false>(s_jf, xf, yf, tf, ioff, F.thi, F.tmaxf, jb, wq_addr, lt, tag, tail);
                    else if (variant == 1)
                        sweep<false, true>(s_jf, xf, yf, tf, ioff, F.thi, F.tmaxf, jb, wq_addr, lt, tag, tail);
                    else
                        sweep<true, true>(s_jf, xf, yf, tf, ioff, F.thi, F.tmaxf, jb, wq_addr, lt, tag, tail);
                }
                drain<MODE, FX>(S, P, ibase, sbase + jbase, wq_addr, lane, lt, head, tail, npass, true);
                npass_total += npass;
                ++items_total;

                // This warp is done with stage st.  The LAST warp to get here refills the stage with
                // the next-but-one item of the run: nobody waits for anybody inside a run.
                __syncwarp();
                if (lane == 0) {
                    // (this warp's reads of the stage have been consumed: its drain is finished.  No
                    // __threadfence_block() here: a fence in the kernel turns every fire-and-forget
                    // REDG of the histogram into an ATOMG that the warp then waits for.)
                    if (atomicAdd(&s_done[st], 1u) == VT / 32 - 1) {
                        s_done[st] = 0u;
                        if (c + NSTAGE < L) {
                            mbar_expect_tx(bar_j + 8 * st, REC_BYTES);
                            bulk_g2s(sbase + jbase, P.recs + (size_t)item_J(c + NSTAGE) * REC_BYTES, REC_BYTES,
                                     bar_j + 8 * st);
                        }
                    }
                }
            }
            items_done += (unsigned)L;
            w += L;
        }
        kq = *s_next;  // written by thread 0 after the barrier of this chunk's first run
    }
    __syncthreads();
    for (int b = tid; b < P.n_bins; b += VT) {
        my_cnt[b] += S.s_cnt[b];
        if (FX) {
            my_limb[b] = S.s_lim[b];
            my_limb[P.n_bins + b] = S.s_lim[P.n_bins + b];
            my_limb[2 * P.n_bins + b] = S.s_lim[2 * P.n_bins + b];
        }
    }
    // stats: pairs inside the window (warp reduce, one atomic per warp)
    for (int off = 16; off > 0; off >>= 1)
        npass_total += __shfl_xor_sync(0xffffffffu, npass_total, off);
    if (lane == 0 && npass_total) atomicAdd(&P.stats[2], npass_total);
    if (tid == 0 && items_total) atomicAdd(&P.stats[3], (unsigned long long)items_total);
}

// Exact signed 128-bit integer -> float64 with ONE round-to-nearest-even.
__device__ __forceinline__ double i128_to_double(__int128 v)
{
    const bool neg = v < 0;
    unsigned __int128 m = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
    const unsigned long long hi = (unsigned long long)(m >> 64), lo = (unsigned long long)m;
    double d;
    if (hi == 0ull) {
        d = __ull2double_rn(lo);
    } else {
        const int s = 64 - __clzll((long long)hi);              // 1 .. 64: keep the top 64 bits
        unsigned long long top = (unsigned long long)(m >> s);
        const unsigned __int128 rest = m & ((((unsigned __int128)1) << s) - 1);
        if (rest != 0) top |= 1ull;                            // sticky bit (64 - 53 guard bits above it)
        d = ldexp(__ull2double_rn(top), s);
    }
    return neg ? -d : d;
}

// Reduction over the per-block partials + the reference's peeled first pair.
//   FX = true: exact 128-bit integer sums of the limb counters, ONE rounding to float64;
//   FX = false: FP64 partial sums added in a fixed order (interleaved slices + a binary tree).
template <bool FX>
__global__ void __launch_bounds__(1024) reduce_kernel(Params P, int nblocks, const double *x, const double *y,
                              const double *t, const double *v, int64_t n, double smax,
                              double ws, double wt, int peel, int counts_f64, unsigned long long *counts,
                              double *sums)
{
    // 1024 threads = 32 bins x 32 slices of the block list: slice s adds the limb counters of
    // blocks k = s, s+32, ... as 128-bit integers (coalesced over the bins), then the 32 slices
    // are added.  Integer sums: the result does not depend on the grouping.
    constexpr int NSL = 32;
    __shared__ unsigned long long s_lo[NSL][32], s_hi[NSL][32], s_cnt8[NSL][32];
    __shared__ unsigned s_poison[NSL][32];
    const int bl = threadIdx.x & 31, ks = threadIdx.x >> 5;
    const int b = blockIdx.x * 32 + bl;
    double s;
    unsigned long long c = 0;
    if (FX) {
        unsigned __int128 a = 0;
        unsigned long long cc = 0;
        unsigned poison = 0;
        if (b < P.n_bins) {
            for (int k = ks; k < nblocks; k += NSL) {
                const unsigned *l = P.partial_limb + (size_t)k * 4 * P.n_bins + b;
                const unsigned c0 = l[0], c1 = l[P.n_bins], c2 = l[2 * P.n_bins], wr = l[3 * P.n_bins];
                a += (unsigned __int128)c0 + ((unsigned __int128)c1 << LIMB_BITS) +
                     ((unsigned __int128)c2 << (2 * LIMB_BITS)) +
                     ((unsigned __int128)(wr & ~POISON) << (2 * LIMB_BITS + 32));
                poison |= wr & POISON;
                cc += P.partial_cnt[(size_t)k * P.n_bins + b];
            }
        }
        s_lo[ks][bl] = (unsigned long long)a;
        s_hi[ks][bl] = (unsigned long long)(a >> 64);
        s_cnt8[ks][bl] = cc;
        s_poison[ks][bl] = poison;
        __syncthreads();
        if (ks != 0 || b >= P.n_bins) return;
        unsigned __int128 tot = 0;
        poison = 0;
#pragma unroll 4
        for (int q = 0; q < NSL; ++q) {
            tot += ((unsigned __int128)s_hi[q][bl] << 64) | (unsigned __int128)s_lo[q][bl];
            c += s_cnt8[q][bl];
            poison |= s_poison[q][bl];
        }
        const FilterConst F = *P.filter;
        // every term carried the bias (covariogram: terms of either sign); remove it exactly
        const __int128 exact = (__int128)tot - (__int128)F.fx_bias * (__int128)c;
        s = i128_to_double(exact) * F.fx_inv;
        if (poison) s = __longlong_as_double(0x7ff8000000000000LL);
    } else {
        double *s_acc = reinterpret_cast<double *>(&s_lo[0][0]);      // [NSL][32]
        double a = 0.0;
        unsigned long long cc = 0;
        if (b < P.n_bins) {
#pragma unroll 8
            for (int k = ks; k < nblocks; k += NSL) {
                a += P.partial_sum[(size_t)k * P.n_bins + b];
                cc += P.partial_cnt[(size_t)k * P.n_bins + b];
            }
        }
        s_acc[ks * 32 + bl] = a;
        s_cnt8[ks][bl] = cc;
        __syncthreads();
        if (ks != 0 || b >= P.n_bins) return;
        double acc[NSL];
#pragma unroll
        for (int q = 0; q < NSL; ++q) {
            acc[q] = s_acc[q * 32 + bl];
            c += s_cnt8[q][bl];
        }
#pragma unroll
        for (int w = 1; w < NSL; w <<= 1)
#pragma unroll
            for (int q = 0; q < NSL; q += 2 * w) acc[q] = acc[q] + acc[q + w];
        s = acc[0];
    }
    if (peel && n >= 2) {
        // rows 0 and 1 of the caller's arrays: the pair kernel binned them with the
        // reciprocal multiply; the reference bins this one pair with a true division.
        const double dx = __dsub_rn(x[0], x[1]), dy = __dsub_rn(y[0], y[1]);
        const double d2 = __fma_rn(dy, dy, __dmul_rn(dx, dx));
        const double lag_s = __dsqrt_rn(d2);
        const double lag_t = fabs(__dsub_rn(t[0], t[1]));
        if (!(lag_s > smax) && !(lag_t > P.tmax) && d2 == d2 && lag_t == lag_t) {
            const long long bs_m = __double2ll_rz(__dmul_rn(lag_s, P.inv_ws));
            const long long bt_m = __double2ll_rz(__dmul_rn(lag_t, P.inv_wt));
            const long long bs_d = __double2ll_rz(__ddiv_rn(lag_s, ws));
            const long long bt_d = __double2ll_rz(__ddiv_rn(lag_t, wt));
            const bool in_m = bs_m >= 0 && bs_m < P.n_space_bins && bt_m >= 0 && bt_m < P.n_time_bins;
            const bool in_d = bs_d >= 0 && bs_d < P.n_space_bins && bt_d >= 0 && bt_d < P.n_time_bins;
            const long long bin_m = in_m ? bs_m * P.n_time_bins + bt_m : -1;
            const long long bin_d = in_d ? bs_d * P.n_time_bins + bt_d : -1;
            if (bin_m != bin_d) {
                double a, bb;
                if (P.mode == 0) {
                    a = __dsub_rn(v[0], v[1]);
                    bb = a;
                } else {
                    a = __dsub_rn(v[0], P.mean);
                    bb = __dsub_rn(v[1], P.mean);
                }
                const double val = __dmul_rn(a, bb);
                if (bin_m == b) {
                    c -= 1;
                    s -= val;
                }
                if (bin_d == b) {
                    c += 1;
                    s += val;
                }
            }
        }
    }
    // counts as float64 (exact below 2^53; the reference's own `norm` array is float64): lets a
    // multi-GPU caller sum counts and sums with ONE all-reduce over one float64 buffer
    if (counts_f64) reinterpret_cast<double *>(counts)[b] = (double)c;
    else counts[b] = c;
    sums[b] = s;
}

// Fixed-order FP64 reduction of per-block partial sums (the sampled-pair branch, whose draws
// are random anyway).
__global__ void __launch_bounds__(1024) reduce_f64_kernel(const double *partial_sum,
                                                          const unsigned long long *partial_cnt, int n_bins,
                                                          int nblocks, unsigned long long *counts, double *sums)
{
    constexpr int NSL = 32;
    __shared__ double s_acc[NSL][32];
    __shared__ unsigned long long s_cnt8[NSL][32];
    const int bl = threadIdx.x & 31, ks = threadIdx.x >> 5;
    const int b = blockIdx.x * 32 + bl;
    {
        double a = 0.0;
        unsigned long long cc = 0;
        if (b < n_bins) {
#pragma unroll 8
            for (int k = ks; k < nblocks; k += NSL) {
                a += partial_sum[(size_t)k * n_bins + b];
                cc += partial_cnt[(size_t)k * n_bins + b];
            }
        }
        s_acc[ks][bl] = a;
        s_cnt8[ks][bl] = cc;
    }
    __syncthreads();
    if (ks != 0 || b >= n_bins) return;
    double acc[NSL];
    unsigned long long c = 0;
#pragma unroll
    for (int q = 0; q < NSL; ++q) {
        acc[q] = s_acc[q][bl];
        c += s_cnt8[q][bl];
    }
#pragma unroll
    for (int w = 1; w < NSL; w <<= 1)
#pragma unroll
        for (int q = 0; q < NSL; q += 2 * w) acc[q] = acc[q] + acc[q + w];
    counts[b] = c;
    sums[b] = acc[0];
}


// ---- sampled-pair branch (el_max / n_samples, stif/utils.py:94-101) -------------------
// n_samples draws of an ordered pair (i, j), each index uniform in [0, n), i == j skipped
// (the draw still counts).  The reference uses numba's OS-seeded MT19937, so only
// statistical parity is possible; here a counter-based Philox4x32-10 stream makes a call
// reproducible for a given seed.
__device__ __forceinline__ void philox4x32_10(unsigned long long ctr, unsigned long long key,
                                              unsigned out[4])
{
    unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
    unsigned k0 = (unsigned)key, k1 = (unsigned)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

struct SampledParams {
    const double *x, *y, *t, *v;
    long long n;
    unsigned long long n_samples, seed;
    double smax, tmax;
};

__global__ void __launch_bounds__(VT) sampled_kernel(Params P, SampledParams Q, double *partial_sum)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_sum = reinterpret_cast<double *>(smem_raw);
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_sum + P.n_bins);
    const int tid = threadIdx.x;
    for (int b = tid; b < P.n_bins; b += VT) {
        s_sum[b] = 0.0;
        s_cnt[b] = 0u;
    }
    unsigned long long *my_cnt = P.partial_cnt + (size_t)blockIdx.x * P.n_bins;
    double *my_sum = partial_sum + (size_t)blockIdx.x * P.n_bins;
    for (int b = tid; b < P.n_bins; b += VT) my_cnt[b] = 0ull;
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * VT;
    unsigned long long done = 0;
    for (unsigned long long base = (unsigned long long)blockIdx.x * VT;; base += stride) {
        // block-uniform loop bounds so the periodic flush barrier is safe
        if (base >= Q.n_samples) break;
        const unsigned long long sidx = base + tid;
        if (sidx < Q.n_samples) {
            unsigned r[4];
            philox4x32_10(sidx, Q.seed, r);
            const unsigned long long ri = ((unsigned long long)r[1] << 32) | r[0];
            const unsigned long long rj = ((unsigned long long)r[3] << 32) | r[2];
            const long long i = (long long)__umul64hi(ri, (unsigned long long)Q.n);
            const long long j = (long long)__umul64hi(rj, (unsigned long long)Q.n);
            if (i != j) {
                const double dx = __dsub_rn(Q.x[i], Q.x[j]), dy = __dsub_rn(Q.y[i], Q.y[j]);
                const double d2 = __fma_rn(dy, dy, __dmul_rn(dx, dx));
                const double lag_s = __dsqrt_rn(d2);
                const double lag_t = fabs(__dsub_rn(Q.t[i], Q.t[j]));
                if (!(lag_s > Q.smax) && !(lag_t > Q.tmax) && d2 == d2 && lag_t == lag_t) {
                    const long long bs = __double2ll_rz(__dmul_rn(lag_s, P.inv_ws));
                    const long long bt = __double2ll_rz(__dmul_rn(lag_t, P.inv_wt));
                    if (bs >= 0 && bs < P.n_space_bins && bt >= 0 && bt < P.n_time_bins) {
                        double a, b;
                        if (P.mode == 0) {
                            a = __dsub_rn(Q.v[i], Q.v[j]);
                            b = a;
                        } else {
                            a = __dsub_rn(Q.v[i], P.mean);
                            b = __dsub_rn(Q.v[j], P.mean);
                        }
                        const int bin = (int)bs * P.n_time_bins + (int)bt;
                        atomicAdd(&s_sum[bin], __dmul_rn(a, b));
                        atomicAdd(&s_cnt[bin], 1u);
                    }
                }
            }
        }
        if (++done == (1ull << 22)) {  // 2^22 x 256 samples per block < 2^32
            __syncthreads();
            for (int b = tid; b < P.n_bins; b += VT) {
                my_cnt[b] += s_cnt[b];
                s_cnt[b] = 0u;
            }
            __syncthreads();
            done = 0;
        }
    }
    __syncthreads();
    for (int b = tid; b < P.n_bins; b += VT) {
        my_cnt[b] += s_cnt[b];
        my_sum[b] = s_sum[b];
    }
}

// Largest double T with sqrt_rn(T) <= smax (host; IEEE sqrt is correctly rounded and
// monotone, so sqrt_rn(d2) > smax  <=>  d2 > T).
// thr[k] (k = 1..ns) = smallest d2 >= 0 with trunc(sqrt_rn(d2) * inv_ws) >= k; thr[0] = -1,
// thr[ns+1] = +inf.  Bisection over the (monotone) bit patterns of non-negative doubles with
// the reference's exact formula, so the device needs no sqrt to find a bin.
static void bin_thresholds(double inv_ws, int ns, double *thr)
{
    auto bin_of = [&](double d2) -> double {
        const double f = sqrt(d2) * inv_ws;
        return (f >= 9.0e18) ? 9.0e18 : (double)(long long)f;
    };
    thr[0] = -1.0;
    thr[ns + 1] = INFINITY;
    for (int k = 1; k <= ns; ++k) {
        unsigned long long lo = 0, hi = 0x7ff0000000000000ull;  // +0 .. +inf
        while (lo < hi) {
            const unsigned long long mid = lo + ((hi - lo) >> 1);
            double d;
            memcpy(&d, &mid, 8);
            if (bin_of(d) >= (double)k) hi = mid; else lo = mid + 1;
        }
        memcpy(&thr[k], &lo, 8);
    }
}

static double sqrt_cut_threshold(double smax)
{
    if (isnan(smax)) return -1.0;                 // "lag > NaN" is never true in IEEE, but
                                                  // fast-math makes it UB; treat as empty
    if (smax < 0.0) return -1.0;                  // sqrt >= 0 > smax for every pair
    if (isinf(smax)) return INFINITY;
    double c = smax * smax;
    if (isinf(c)) c = 1.79769313486231570815e308;
    while (sqrt(c) > smax) c = nextafter(c, -INFINITY);
    for (;;) {
        double up = nextafter(c, INFINITY);
        if (isinf(up) || sqrt(up) > smax) break;
        c = up;
    }
    return c;
}

}  // namespace vario
}  // namespace stif

using namespace stif;
using namespace stif::vario;

extern "C" int stif_variogram_dev(stif_ctx *c, const double *x, const double *y, const double *t,
                                  const double *v, int64_t n, double smax, double tmax, int ns,
                                  int nt, int mode, double mean, int flags, int part, int nparts,
                                  uint64_t *counts_dev, double *sums_dev, void *stream_v)
{
    if (!c) {
        set_error("stif_variogram: ctx is NULL");
        return STIF_EINVAL;
    }
    if (ns <= 0 || nt <= 0 || (int64_t)ns * nt > MAX_BINS) {
        set_error("stif_variogram: n_space_bins*n_time_bins must be in [1, %d] (got %d x %d)",
                  MAX_BINS, ns, nt);
        return STIF_EINVAL;
    }
    if (n < 0 || n > 0x7fffffff - 2 * TI) {
        set_error("stif_variogram: n = %lld out of range", (long long)n);
        return STIF_EINVAL;
    }
    if (mode != 0 && mode != 1) {
        set_error("stif_variogram: mode must be 0 (variogram) or 1 (covariogram)");
        return STIF_EINVAL;
    }
    if (nparts < 1 || part < 0 || part >= nparts) {
        set_error("stif_variogram: bad partition %d of %d", part, nparts);
        return STIF_EINVAL;
    }
    if (!counts_dev || !sums_dev || (n > 0 && (!x || !y || !t || !v))) {
        set_error("stif_variogram: NULL pointer argument");
        return STIF_EINVAL;
    }
    STIF_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : c->stream;
    c->last_stream = st;
    const int n_bins = ns * nt;
    const bool cull = !(flags & STIF_VARIO_NO_CULL);

    // [0..3] counters, [4..11] bounding box of (x, y, t, v) (bbox_kernel), then the FilterConst
    STIF_TRY(c->v_stats.reserve(12 * sizeof(unsigned long long) + sizeof(FilterConst)));
    STIF_CUDA(cudaMemsetAsync(c->v_stats.p, 0, 12 * sizeof(unsigned long long), st));
    STIF_CUDA(cudaEventRecord(c->ev[0], st));
    c->n_phase = 3;

    if (n < 2) {
        STIF_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(uint64_t) * n_bins, st));
        STIF_CUDA(cudaMemsetAsync(sums_dev, 0, sizeof(double) * n_bins, st));
        for (int i = 1; i <= 3; ++i) STIF_CUDA(cudaEventRecord(c->ev[i], st));
        c->v_stats_host[0] = c->v_stats_host[1] = c->v_stats_host[3] = 0;
        return STIF_OK;
    }

    const int64_t n_pad = (n + TI - 1) / TI * TI;
    const int nI = (int)(n_pad / TI), nJ = (int)(n_pad / TJ);
    // space-time sort mode: pays once a time slab holds many sub-tiles (auto: n >= 160 x 1024,
    // measured break-even incl. the longer prepare phase ~150k at the C2 / C3 densities); the
    // explicit item list is sized by its upper bound (every J >= I * JPI), 4 bytes per item
    const int64_t items_max = (int64_t)nI * nJ - (int64_t)JPI * nI * (nI - 1) / 2;
    bool spatial = cull && !(flags & STIF_VARIO_TIME_SORT) &&
                   ((flags & STIF_VARIO_SPACE_SORT) || n >= 160 * 1024);
    if (nJ > 65535 || items_max > (1ll << 27)) spatial = false;
    unsigned long long *ext = c->v_stats.as<unsigned long long>() + 4;
    FilterConst *filter = reinterpret_cast<FilterConst *>(ext + 8);

    // 1. sort indices by time (or by time slab and space cell)
    STIF_TRY(c->v_idx.reserve(sizeof(unsigned) * n));
    STIF_TRY(c->v_idx_out.reserve(sizeof(unsigned) * n));
    const int tb = 256;
    const unsigned nb_n = (unsigned)((n + tb - 1) / tb);
    const unsigned *order = nullptr;
    if (spatial) {
        if (c->v_wait_upload)     // the keys need x and y as well (v still travels under the sort)
            STIF_CUDA(cudaStreamWaitEvent(st, c->v_upload_xy_ev, 0));
        STIF_TRY(c->v_keys.reserve(sizeof(double) * n));
        STIF_TRY(c->v_keys_out.reserve(sizeof(double) * n));
        const double slab_w = fabs(tmax) * (getenv("STIF_VARIO_SLAB") ? atof(getenv("STIF_VARIO_SLAB")) : 0.5);
        bbox_kernel<<<c->sm_count * 4, tb, 0, st>>>(x, y, t, nullptr, n, ext);
        spacetime_keys_kernel<<<nb_n, tb, 0, st>>>(x, y, t, n, ext, slab_w, c->v_keys.as<unsigned>(),
                                                   c->v_idx.as<unsigned>());
        size_t tmp_bytes = 0;
        STIF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, c->v_keys.as<unsigned>(),
                                                  c->v_keys_out.as<unsigned>(),
                                                  c->v_idx.as<unsigned>(),
                                                  c->v_idx_out.as<unsigned>(), (int)n, 0, 32, st));
        STIF_TRY(c->v_cubtmp.reserve(tmp_bytes));
        STIF_CUDA(cub::DeviceRadixSort::SortPairs(c->v_cubtmp.p, tmp_bytes, c->v_keys.as<unsigned>(),
                                                  c->v_keys_out.as<unsigned>(),
                                                  c->v_idx.as<unsigned>(),
                                                  c->v_idx_out.as<unsigned>(), (int)n, 0, 32, st));
        order = c->v_idx_out.as<unsigned>();
    } else if (cull) {
        STIF_TRY(c->v_keys.reserve(sizeof(double) * n));
        STIF_TRY(c->v_keys_out.reserve(sizeof(double) * n));
        pack_keys_kernel<<<nb_n, tb, 0, st>>>(t, n, c->v_keys.as<double>(), c->v_idx.as<unsigned>());
        size_t tmp_bytes = 0;
        STIF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, c->v_keys.as<double>(),
                                                  c->v_keys_out.as<double>(),
                                                  c->v_idx.as<unsigned>(),
                                                  c->v_idx_out.as<unsigned>(), (int)n, 32, 64, st));
        STIF_TRY(c->v_cubtmp.reserve(tmp_bytes));
        STIF_CUDA(cub::DeviceRadixSort::SortPairs(c->v_cubtmp.p, tmp_bytes, c->v_keys.as<double>(),
                                                  c->v_keys_out.as<double>(),
                                                  c->v_idx.as<unsigned>(),
                                                  c->v_idx_out.as<unsigned>(), (int)n, 32, 64, st));
        order = c->v_idx_out.as<unsigned>();
    } else {
        iota_kernel<<<nb_n, tb, 0, st>>>(n, c->v_idx.as<unsigned>());
        order = c->v_idx.as<unsigned>();
    }
    // 2. bounding box -> pre-filter / fixed-point constants; gather into sorted, padded records
    STIF_TRY(c->v_pts.reserve((size_t)REC_BYTES * (size_t)nJ));
    if (c->v_wait_upload) {   // host entry point: x, y, v arrive on the copy stream under the sort
        STIF_CUDA(cudaStreamWaitEvent(st, c->v_upload_ev, 0));
        c->v_wait_upload = false;
    }
    if (spatial) bbox_kernel<<<c->sm_count * 4, tb, 0, st>>>(nullptr, nullptr, nullptr, v, n, ext);
    else bbox_kernel<<<c->sm_count * 4, tb, 0, st>>>(x, y, t, v, n, ext);
    filter_setup_kernel<<<1, 1, 0, st>>>(ext, smax, tmax, mode, mean, filter);
    unsigned char *recs = c->v_pts.as<unsigned char>();
    gather_kernel<<<(unsigned)((n_pad + tb - 1) / tb), tb, 0, st>>>(x, y, t, v, order, n, n_pad, recs, filter);
    // 3. work list
    STIF_TRY(c->v_tileinfo.reserve(sizeof(int) * (3 * (size_t)nI + 1)));
    int *count = c->v_tileinfo.as<int>();
    int *jbeg = count + nI;
    int *prefix = jbeg + nI;
    const double tcut = sqrt_cut_threshold(smax);
    unsigned *items = nullptr;
    double *bb = nullptr;
    if (spatial) {
        STIF_TRY(c->v_bb.reserve(sizeof(double) * 6 * (size_t)nJ));
        STIF_TRY(c->v_items.reserve(sizeof(unsigned) * (size_t)items_max));
        bb = c->v_bb.as<double>();
        items = c->v_items.as<unsigned>();
        tile_bbox_kernel<<<(nJ + 3) / 4, 128, 0, st>>>(recs, nJ, bb);
        items_kernel<false><<<(nI + 3) / 4, 128, 0, st>>>(bb, nI, nJ, tcut, tmax, count, nullptr, nullptr);
        scan_kernel<<<1, 1024, 0, st>>>(count, nI, prefix);
        items_kernel<true><<<(nI + 3) / 4, 128, 0, st>>>(bb, nI, nJ, tcut, tmax, nullptr, prefix, items);
    } else {
        tile_range_kernel<<<(nI + 127) / 128, 128, 0, st>>>(recs, n, nI, nJ, tmax, cull ? 1 : 0, count, jbeg);
        scan_kernel<<<1, 1024, 0, st>>>(count, nI, prefix);
    }
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaEventRecord(c->ev[1], st));

    // 4. pair kernel
    Params P;
    P.recs = recs;
    P.item_prefix = prefix;
    P.item_jbeg = jbeg;
    P.nI = nI;
    P.n_bins = n_bins;
    P.n_time_bins = nt;
    P.n_space_bins = ns;
    P.mode = mode;
    P.part = part;
    P.nparts = nparts;
    P.tcut = tcut;
    P.items = items;
    P.bb = bb;
    P.tmax = tmax;
    const double ws = smax / (double)ns, wt = tmax / (double)nt;  // utils.py:146-147
    P.inv_ws = 1.0 / ws;
    P.inv_wt = 1.0 / wt;
    P.mean = mean;
    P.stats = c->v_stats.as<unsigned long long>();
    P.inv_ws_f = (float)P.inv_ws;
    P.guard = (float)ns * 1.9073486328125e-06f;  // 2^-19 * ns (ns <= 8192: < 0.016)
    P.sorted = (cull && !spatial) ? 1 : 0;
    P.filter = filter;
    P.use_thr = (P.inv_ws > 0.0 && isfinite(P.inv_ws) && isfinite((double)P.inv_ws_f)) ? 1 : 0;
    if (!P.use_thr) P.guard = 2.0f;  // no pair trusts the float guess
    P.guard_hi = 1.0f - P.guard;
    P.ns_f = (float)ns;
    {
        // context-owned host copy: stays valid until the next call, so no synchronisation
        std::vector<double> &thr = c->v_thr_host;
        thr.assign((size_t)ns + 2, 0.0);
        if (P.use_thr) bin_thresholds(P.inv_ws, ns, thr.data());
        STIF_TRY(c->v_thr.reserve(sizeof(double) * thr.size()));
        STIF_CUDA(cudaMemcpyAsync(c->v_thr.p, thr.data(), sizeof(double) * thr.size(),
                                  cudaMemcpyHostToDevice, st));
        P.thr = c->v_thr.as<double>();
    }

    const bool fx = (flags & STIF_VARIO_DETERMINISTIC) != 0;
    void (*kern)(Params) = fx ? (mode == 0 ? pair_kernel<0, true> : pair_kernel<1, true>)
                              : (mode == 0 ? pair_kernel<0, false> : pair_kernel<1, false>);
    const size_t smem = O_THR + sizeof(double) * ((size_t)ns + 2) +
                        (fx ? 4 : 1) * sizeof(unsigned) * (size_t)n_bins;
    {
        int smem_max = 0;
        STIF_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
        if (smem > (size_t)smem_max) {
            set_error("stif_variogram: %d x %d bins need %zu B of shared memory per block (limit %d)%s", ns, nt,
                      smem, smem_max, fx ? "; STIF_VARIO_DETERMINISTIC keeps four histograms per block -- use "
                      "fewer bins or the default mode" : "");
            return STIF_EINVAL;
        }
    }
    STIF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    int occ = 0;
    STIF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, VT, smem));
    if (occ < 1) {
        set_error("stif_variogram: pair kernel does not fit (smem %zu B)", smem);
        return STIF_ECUDA;
    }
    const int grid = c->sm_count * occ;
    STIF_TRY(c->v_partial.reserve((4 * sizeof(unsigned) + sizeof(unsigned long long)) * (size_t)grid * n_bins));
    P.partial_cnt = c->v_partial.as<unsigned long long>();
    P.partial_limb = reinterpret_cast<unsigned *>(P.partial_cnt + (size_t)grid * n_bins);
    P.partial_sum = reinterpret_cast<double *>(P.partial_limb);     // the two modes share the space
    kern<<<grid, VT, smem, st>>>(P);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaEventRecord(c->ev[2], st));

    // 5. reduce (+ peel fix-up on partition 0)
    const int peel = ((flags & STIF_VARIO_PEEL) && part == 0) ? 1 : 0;
    if (fx)
        reduce_kernel<true><<<(n_bins + 31) / 32, 1024, 0, st>>>(
            P, grid, x, y, t, v, n, smax, ws, wt, peel, (flags & STIF_VARIO_COUNTS_F64) ? 1 : 0,
            reinterpret_cast<unsigned long long *>(counts_dev), sums_dev);
    else
        reduce_kernel<false><<<(n_bins + 31) / 32, 1024, 0, st>>>(
            P, grid, x, y, t, v, n, smax, ws, wt, peel, (flags & STIF_VARIO_COUNTS_F64) ? 1 : 0,
            reinterpret_cast<unsigned long long *>(counts_dev), sums_dev);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaEventRecord(c->ev[3], st));

    const uint64_t all_pairs = (uint64_t)n * (uint64_t)(n - 1) / 2;
    c->v_stats_host[0] = all_pairs / (uint64_t)nparts;
    return STIF_OK;
}

extern "C" int stif_variogram_stats(stif_ctx *c, uint64_t out[4])
{
    if (!c || !out) {
        set_error("stif_variogram_stats: NULL argument");
        return STIF_EINVAL;
    }
    STIF_CUDA(cudaSetDevice(c->device));
    out[0] = c->v_stats_host[0];
    out[1] = out[2] = out[3] = 0;
    if (!c->v_stats.p || !c->v_tileinfo.p) return STIF_OK;
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    STIF_CUDA(cudaStreamSynchronize(st));
    unsigned long long h[4];
    STIF_CUDA(cudaMemcpy(h, c->v_stats.p, sizeof(h), cudaMemcpyDeviceToHost));
    out[2] = h[2];
    out[3] = h[3];
    out[1] = h[3] * (uint64_t)TI * TJ;
    return STIF_OK;
}

extern "C" int stif_variogram_host(stif_ctx *c, const double *x, const double *y, const double *t,
                                   const double *v, int64_t n, double smax, double tmax, int ns,
                                   int nt, int mode, double mean, int flags, int part, int nparts,
                                   uint64_t *counts_host, double *sums_host)
{
    if (!c) {
        set_error("stif_variogram: ctx is NULL");
        return STIF_EINVAL;
    }
    if (ns <= 0 || nt <= 0 || (int64_t)ns * nt > MAX_BINS || n < 0) {
        set_error("stif_variogram: bad sizes (n=%lld, bins %d x %d, limit %d bins)", (long long)n,
                  ns, nt, MAX_BINS);
        return STIF_EINVAL;
    }
    STIF_CUDA(cudaSetDevice(c->device));
    const size_t nb = (size_t)ns * nt;
    const size_t in_bytes = sizeof(double) * (size_t)n;
    const size_t out_off = (4 * in_bytes + 255) / 256 * 256;
    STIF_TRY(c->v_stage.reserve(out_off + nb * 16));
    char *base = c->v_stage.as<char>();
    double *dx = (double *)base, *dy = dx + n, *dt = dy + n, *dv = dt + n;
    uint64_t *dc = (uint64_t *)(base + out_off);
    double *ds = (double *)(base + out_off + nb * 8);
    cudaStream_t st = c->stream;
    if (n > 0) {
        // the time sort needs t only: x, y, v travel on the copy stream underneath it and the
        // gather waits for them (an event); the previous call's kernels are done with the
        // staging buffers because every host call ends with a stream synchronisation
        if (!c->stream2) STIF_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        if (!c->v_upload_ev) STIF_CUDA(cudaEventCreateWithFlags(&c->v_upload_ev, cudaEventDisableTiming));
        if (!c->v_upload_xy_ev) STIF_CUDA(cudaEventCreateWithFlags(&c->v_upload_xy_ev, cudaEventDisableTiming));
        STIF_CUDA(cudaMemcpyAsync(dt, t, in_bytes, cudaMemcpyHostToDevice, st));
        STIF_CUDA(cudaMemcpyAsync(dx, x, in_bytes, cudaMemcpyHostToDevice, c->stream2));
        STIF_CUDA(cudaMemcpyAsync(dy, y, in_bytes, cudaMemcpyHostToDevice, c->stream2));
        STIF_CUDA(cudaEventRecord(c->v_upload_xy_ev, c->stream2));
        STIF_CUDA(cudaMemcpyAsync(dv, v, in_bytes, cudaMemcpyHostToDevice, c->stream2));
        STIF_CUDA(cudaEventRecord(c->v_upload_ev, c->stream2));
        c->v_wait_upload = true;
    }
    const int rc = stif_variogram_dev(c, dx, dy, dt, dv, n, smax, tmax, ns, nt, mode, mean, flags,
                                      part, nparts, dc, ds, st);
    if (c->v_wait_upload) {   // early return before the gather: still join the copy stream
        c->v_wait_upload = false;
        STIF_CUDA(cudaStreamWaitEvent(st, c->v_upload_ev, 0));
    }
    if (rc != STIF_OK) {
        cudaStreamSynchronize(st);   // the caller may free its arrays
        return rc;
    }
    STIF_CUDA(cudaMemcpyAsync(counts_host, dc, nb * 8, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaMemcpyAsync(sums_host, ds, nb * 8, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaStreamSynchronize(st));
    return STIF_OK;
}

namespace stif {
namespace vario {
__global__ void select_obs_kernel(const double *ox, const double *oy, const double *ot, const double *ov,
                                  const long long *idx, int64_t n_idx, int64_t n_obs, double *x, double *y,
                                  double *t, double *v, int *bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_idx) return;
    long long k = idx[i];
    if (k < 0) k += n_obs;                       // numpy-style negative indices
    if (k < 0 || k >= n_obs) {
        *bad = 1;
        k = 0;
    }
    x[i] = ox[k];
    y[i] = oy[k];
    t[i] = ot[k];
    v[i] = ov[k];
}
}  // namespace vario
}  // namespace stif

// Variogram / covariogram of the BOUND observation set (stif_kriging_set_obs: coordinates and
// residuals already on the device), optionally restricted to an index list.  This is what the
// cross-validation driver needs per fold (stif/predictor.py:293-296 calc_empirical_variogram(train)):
// the coordinates never travel again, a fold that refits the covariate model replaces only the
// residuals (stif_kriging_set_residuals), and a subset costs 8 B per selected observation (its
// index) instead of 32 B.
extern "C" int stif_variogram_bound(stif_ctx *c, const int64_t *idx_host, int64_t n_idx, double smax,
                                    double tmax, int ns, int nt, int mode, double mean, int flags, int part,
                                    int nparts, uint64_t *counts_host, double *sums_host)
{
    if (!c) {
        set_error("stif_variogram_bound: ctx is NULL");
        return STIF_EINVAL;
    }
    if (c->n_obs <= 0 || !c->k_obs.p) {
        set_error("stif_variogram_bound: no observations bound (stif_kriging_set_obs)");
        return STIF_ESTATE;
    }
    if (ns <= 0 || nt <= 0 || (int64_t)ns * nt > MAX_BINS || n_idx < 0 || !counts_host || !sums_host) {
        set_error("stif_variogram_bound: bad arguments (n_idx=%lld, bins %d x %d)", (long long)n_idx, ns, nt);
        return STIF_EINVAL;
    }
    STIF_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int64_t n_obs = c->n_obs;
    const double *ox = c->k_obs.as<double>(), *oy = ox + n_obs, *ot = oy + n_obs, *ov = ot + n_obs;
    const size_t nb = (size_t)ns * nt;
    const int64_t n = idx_host ? n_idx : n_obs;
    const size_t arr = (sizeof(double) * (size_t)n + 255) / 256 * 256;
    // staging: [counts | sums | bad flag] then, for a subset, x y t v and the index list
    STIF_TRY(c->v_stage.reserve(nb * 16 + 256 + (idx_host ? 5 * arr : 0)));
    char *base = c->v_stage.as<char>();
    uint64_t *dc = (uint64_t *)base;
    double *ds = (double *)(base + nb * 8);
    int *dbad = (int *)(base + nb * 16);
    const double *x = ox, *y = oy, *t = ot, *v = ov;
    if (idx_host) {
        char *p = base + nb * 16 + 256;
        double *sx = (double *)p, *sy = (double *)(p + arr), *stt = (double *)(p + 2 * arr), *sv = (double *)(p + 3 * arr);
        long long *didx = (long long *)(p + 4 * arr);
        STIF_CUDA(cudaMemsetAsync(dbad, 0, sizeof(int), st));
        if (n > 0) {
            STIF_CUDA(cudaMemcpyAsync(didx, idx_host, sizeof(long long) * (size_t)n, cudaMemcpyHostToDevice, st));
            select_obs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ox, oy, ot, ov, didx, n, n_obs, sx, sy, stt,
                                                                          sv, dbad);
            STIF_CUDA(cudaGetLastError());
        }
        x = sx; y = sy; t = stt; v = sv;
    }
    const int rc = stif_variogram_dev(c, x, y, t, v, n, smax, tmax, ns, nt, mode, mean, flags, part, nparts, dc, ds, st);
    if (rc != STIF_OK) {
        cudaStreamSynchronize(st);
        return rc;
    }
    int bad = 0;
    STIF_CUDA(cudaMemcpyAsync(counts_host, dc, nb * 8, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaMemcpyAsync(sums_host, ds, nb * 8, cudaMemcpyDeviceToHost, st));
    if (idx_host) STIF_CUDA(cudaMemcpyAsync(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaStreamSynchronize(st));
    if (bad) {
        set_error("stif_variogram_bound: index out of range for %lld bound observations", (long long)n_obs);
        return STIF_EINVAL;
    }
    return STIF_OK;
}

extern "C" int stif_variogram_sampled_dev(stif_ctx *c, const double *x, const double *y,
                                          const double *t, const double *v, int64_t n, double smax,
                                          double tmax, int ns, int nt, int mode, double mean,
                                          uint64_t n_samples, uint64_t seed, uint64_t *counts_dev,
                                          double *sums_dev, void *stream_v)
{
    if (!c) {
        set_error("stif_variogram_sampled: ctx is NULL");
        return STIF_EINVAL;
    }
    if (ns <= 0 || nt <= 0 || (int64_t)ns * nt > MAX_BINS || n < 0 || (mode != 0 && mode != 1) ||
        !counts_dev || !sums_dev || (n > 0 && (!x || !y || !t || !v))) {
        set_error("stif_variogram_sampled: bad arguments (n=%lld, bins %d x %d)", (long long)n, ns, nt);
        return STIF_EINVAL;
    }
    STIF_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : c->stream;
    c->last_stream = st;
    const int n_bins = ns * nt;
    STIF_CUDA(cudaEventRecord(c->ev[0], st));
    STIF_CUDA(cudaEventRecord(c->ev[1], st));
    c->n_phase = 3;
    if (n < 2 || n_samples == 0) {
        STIF_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(uint64_t) * n_bins, st));
        STIF_CUDA(cudaMemsetAsync(sums_dev, 0, sizeof(double) * n_bins, st));
        STIF_CUDA(cudaEventRecord(c->ev[2], st));
        STIF_CUDA(cudaEventRecord(c->ev[3], st));
        return STIF_OK;
    }
    Params P;
    memset(&P, 0, sizeof(P));
    P.n_bins = n_bins;
    P.n_time_bins = nt;
    P.n_space_bins = ns;
    P.mode = mode;
    P.tmax = tmax;
    const double ws = smax / (double)ns, wt = tmax / (double)nt;
    P.inv_ws = 1.0 / ws;
    P.inv_wt = 1.0 / wt;
    P.mean = mean;
    SampledParams Q;
    Q.x = x;
    Q.y = y;
    Q.t = t;
    Q.v = v;
    Q.n = n;
    Q.n_samples = n_samples;
    Q.seed = seed;
    Q.smax = smax;
    Q.tmax = tmax;
    const size_t smem = (sizeof(double) + sizeof(unsigned)) * (size_t)n_bins;
    STIF_CUDA(cudaFuncSetAttribute(sampled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    int grid = c->sm_count * 4;
    const unsigned long long want = (n_samples + VT - 1) / VT;
    if ((unsigned long long)grid > want) grid = (int)want;
    STIF_TRY(c->v_partial.reserve((sizeof(double) + sizeof(unsigned long long)) * (size_t)grid * n_bins));
    double *partial_sum = c->v_partial.as<double>();
    P.partial_cnt = reinterpret_cast<unsigned long long *>(partial_sum + (size_t)grid * n_bins);
    sampled_kernel<<<grid, VT, smem, st>>>(P, Q, partial_sum);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaEventRecord(c->ev[2], st));
    reduce_f64_kernel<<<(n_bins + 31) / 32, 1024, 0, st>>>(
        partial_sum, P.partial_cnt, n_bins, grid, reinterpret_cast<unsigned long long *>(counts_dev), sums_dev);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaEventRecord(c->ev[3], st));
    return STIF_OK;
}

extern "C" int stif_variogram_sampled_host(stif_ctx *c, const double *x, const double *y,
                                           const double *t, const double *v, int64_t n, double smax,
                                           double tmax, int ns, int nt, int mode, double mean,
                                           uint64_t n_samples, uint64_t seed, uint64_t *counts_host,
                                           double *sums_host)
{
    if (!c) {
        set_error("stif_variogram_sampled: ctx is NULL");
        return STIF_EINVAL;
    }
    if (ns <= 0 || nt <= 0 || (int64_t)ns * nt > MAX_BINS || n < 0) {
        set_error("stif_variogram_sampled: bad sizes (n=%lld, bins %d x %d)", (long long)n, ns, nt);
        return STIF_EINVAL;
    }
    STIF_CUDA(cudaSetDevice(c->device));
    const size_t nb = (size_t)ns * nt;
    const size_t in_bytes = sizeof(double) * (size_t)n;
    const size_t out_off = (4 * in_bytes + 255) / 256 * 256;
    STIF_TRY(c->v_stage.reserve(out_off + nb * 16));
    char *base = c->v_stage.as<char>();
    double *dx = (double *)base, *dy = dx + n, *dt = dy + n, *dv = dt + n;
    uint64_t *dc = (uint64_t *)(base + out_off);
    double *ds = (double *)(base + out_off + nb * 8);
    cudaStream_t st = c->stream;
    if (n > 0) {
        STIF_CUDA(cudaMemcpyAsync(dx, x, in_bytes, cudaMemcpyHostToDevice, st));
        STIF_CUDA(cudaMemcpyAsync(dy, y, in_bytes, cudaMemcpyHostToDevice, st));
        STIF_CUDA(cudaMemcpyAsync(dt, t, in_bytes, cudaMemcpyHostToDevice, st));
        STIF_CUDA(cudaMemcpyAsync(dv, v, in_bytes, cudaMemcpyHostToDevice, st));
    }
    STIF_TRY(stif_variogram_sampled_dev(c, dx, dy, dt, dv, n, smax, tmax, ns, nt, mode, mean,
                                        n_samples, seed, dc, ds, st));
    STIF_CUDA(cudaMemcpyAsync(counts_host, dc, nb * 8, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaMemcpyAsync(sums_host, ds, nb * 8, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaStreamSynchronize(st));
    return STIF_OK;
}
