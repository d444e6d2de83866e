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
//   1. pack time keys, radix-sort observation indices by time (CUB; plumbing)
//   2. gather into a time-sorted AoS double4 {x,y,t,v} array, NaN-padded to a tile multiple
//   3. per i-tile (1024 obs) find the last j-sub-tile (256 obs) that can still be inside
//      the time window -> work list of (i-tile, j-sub-tile) items; every other tile pair
//      is provably outside |dt| <= tmax and is never visited ("culling")
//   4. persistent pair kernel: grid = SMs x resident CTAs; block b takes items
//      b, b+grid, ... of its partition (static, deterministic).  Per item each thread
//      keeps 4 i-points in registers and sweeps the 256 j-points broadcast from shared
//      memory: 7 FP64-pipe instructions per pair.  Pairs inside the window are pushed to
//      a per-thread shared-memory queue (2 predicated instructions, no branch) and binned
//      later in a dense drain loop with shared-memory histogram atomics.
//   5. deterministic fixed-order reduction of the per-block partial histograms.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <vector>

namespace stif {
namespace vario {

constexpr int VT = 256;        // threads per block
constexpr int VR = 4;          // i-points per thread (registers)
constexpr int TI = VT * VR;    // i-tile: 1024 observations
constexpr int TJ = 256;        // j-sub-tile: 256 observations (8-bit index in the queue)
constexpr int JPI = TI / TJ;   // j-sub-tiles per i-tile
constexpr int QCAP = 32;       // queue slots per thread (uint16 each)
constexpr int VG = 4;          // j-iterations between queue-overflow checks
constexpr int MAX_BINS = 8192; // shared-memory histogram limit (12 B per bin)
constexpr unsigned FLUSH_ITEMS = 8192;  // 2^13 items x 2^18 pairs < 2^32: u32 counts are safe

struct Params {
    const double4 *pts;   // sorted, padded observations
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
    double *partial_sum;           // [grid][n_bins]
    unsigned long long *partial_cnt;  // [grid][n_bins]
    unsigned long long *stats;     // [4]
    const double *thr;    // [n_space_bins + 2] d2 thresholds: bin b <=> thr[b] <= d2 < thr[b+1]
    int use_thr;          // 0: evaluate sqrt/mul/trunc per binned pair (degenerate widths)
    float inv_ws_f;       // float copy of inv_ws for the first guess
};

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

__global__ void gather_kernel(const double *x, const double *y, const double *t, const double *v,
                              const unsigned *idx, int64_t n, int64_t n_pad, double4 *pts)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pad) return;
    double4 p;
    if (k < n) {
        unsigned i = idx[k];
        p = make_double4(x[i], y[i], t[i], v[i]);
    } else {
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        p = make_double4(qnan, qnan, qnan, 0.0);  // NaN never passes d2 <= tcut
    }
    pts[k] = p;
}

// One thread per i-tile: last j-sub-tile that may hold a pair with rn(tj - ti) <= tmax.
// Sorted ascending by t, so the smallest lag between i-tile I and sub-tile J (beyond the
// diagonal) is rn(t[J*TJ] - t_last(I)), and rounding is monotone.
__global__ void tile_range_kernel(const double4 *pts, int64_t n, int nI, int nJ, double tmax,
                                  int cull, int *count, int *jbeg)
{
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nI) return;
    const int jb = I * JPI;
    int jend = nJ - 1;
    if (cull) {
        int64_t last = (int64_t)I * TI + TI - 1;
        if (last > n - 1) last = n - 1;
        const double ti_max = pts[last].z;
        // first J in (jb+JPI-1, nJ) with lag > tmax; all later ones are also outside
        int lo = jb + JPI, hi = nJ;  // search in [lo, hi)
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            double lag = __dsub_rn(pts[(int64_t)mid * TJ].z, ti_max);
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

struct Smem {
    double *g_sum;         // [n_bins] this block's partial sums in GLOBAL memory (native L2 f64 atomics)
    double4 *s_i;          // [TI]
    double4 *s_j;          // [TJ]
    unsigned short *q;     // [QCAP][VT]
    unsigned *s_cnt;       // [n_bins]
    double *s_thr;         // [n_space_bins + 2]
    unsigned *stg;         // [VT/32][64] per-warp staging of compacted queue entries
};

__device__ __forceinline__ void st_shared_u16(unsigned addr, unsigned short v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

// Histogram update of one binned pair.  The count is a native shared-memory atomic
// (ATOMS.POPC.INC).  The FP64 sum goes to the block's partial histogram in GLOBAL memory with a
// native L2 atomic (REDG.E.ADD.F64, fire and forget): on sm_100a a shared-memory f64 add is a
// compare-and-swap loop (ATOMS.CAST.SPIN.64, 2.95 tries per warp on C2 because a warp's pairs
// share the time bin) and was 16 % of the kernel's stall samples; the block's 7.2 KB of
// partial sums stay in L2.  A warp-aggregated
// variant (match_any + shuffle reduction per distinct bin) was measured 2x SLOWER on C2
// (profiles/r01_notes.md): a warp's 32 staged pairs spread over ~25 distinct bins, so there
// is little to aggregate and the extra shuffles dominate.
__device__ __forceinline__ void accumulate(const Smem &S, int bin /* < 0: nothing */, double val)
{
    if (bin >= 0) {
        atomicAdd(&S.g_sum[bin], val);
        atomicAdd(&S.s_cnt[bin], 1u);
    }
}

// Bin one in-window pair per lane.  e = (owner lane << 16) | (r << 8) | j.
__device__ __forceinline__ void bin_entry(const Smem &S, const Params &P, int warp_base, int lane,
                                          unsigned e, bool have)
{
    int bin = -1;
    double val = 0.0;
    if (have) {
        const int j = e & 0xff, r = (e >> 8) & 0xff, owner = e >> 16;
        const double4 pi = S.s_i[r * VT + warp_base + owner];
        const double4 pj = S.s_j[j];
        const double dx = __dsub_rn(pi.x, pj.x);
        const double dy = __dsub_rn(pi.y, pj.y);
        const double d2 = __fma_rn(dy, dy, __dmul_rn(dx, dx));
        const double lag_t = fabs(__dsub_rn(pi.z, pj.z));
        // both bin indices are bounded by the cut-offs that let the pair in (lag <= max =
        // n_bins * width), so 32-bit conversions are exact
        int bs;
        if (P.use_thr) {
            // trunc(sqrt_rn(d2)*inv_ws) is monotone in d2: look the bin up in the exact d2
            // thresholds (host-bisected with the reference's formula) from a float guess
            float f = __double2float_rn(d2), sq;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(f));
            int b = __float2int_rz(fminf(sq * P.inv_ws_f, (float)P.n_space_bins));
            b = b < 0 ? 0 : b;
            while (d2 < S.s_thr[b]) --b;
            while (d2 >= S.s_thr[b + 1]) ++b;
            bs = b;
        } else {
            bs = __double2int_rz(__dmul_rn(__dsqrt_rn(d2), P.inv_ws));
        }
        const int bt = __double2int_rz(__dmul_rn(lag_t, P.inv_wt));
        if ((unsigned)bs < (unsigned)P.n_space_bins && (unsigned)bt < (unsigned)P.n_time_bins) {
            if (P.mode == 0) {
                const double a = __dsub_rn(pi.w, pj.w);
                val = __dmul_rn(a, a);
            } else {
                val = __dmul_rn(__dsub_rn(pi.w, P.mean), __dsub_rn(pj.w, P.mean));
            }
            bin = bs * P.n_time_bins + bt;
        }
    }
    accumulate(S, bin, val);
}

// Drain the 32 per-thread queues of a warp.  Slot s of every lane that has one is compacted
// (ballot) into the warp's staging buffer; whenever 32 entries are staged they are binned
// with all lanes active.  `nst` (warp-uniform) carries the < 32 leftovers to the next call;
// `flush` bins them too (end of an item, before the tiles are replaced).
__device__ __forceinline__ void drain(const Smem &S, const Params &P, int tid, int cnt, int &nst,
                                      bool flush)
{
    const int lane = tid & 31, warp_base = tid & ~31;
    unsigned *stg = S.stg + (tid >> 5) * 64;
    const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
    for (int s = 0; s < maxcnt; ++s) {
        const bool has = s < cnt;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (has) stg[nst + __popc(bal & ((1u << lane) - 1u))] = ((unsigned)lane << 16) | S.q[s * VT + tid];
        nst += __popc(bal);
        __syncwarp();
        if (nst >= 32) {
            bin_entry(S, P, warp_base, lane, stg[lane], true);
            const int rem = nst - 32;
            const unsigned mv = (lane < rem) ? stg[32 + lane] : 0u;
            __syncwarp();
            if (lane < rem) stg[lane] = mv;
            nst = rem;
            __syncwarp();
        }
    }
    if (flush && nst > 0) {
        bin_entry(S, P, warp_base, lane, stg[lane], lane < nst);
        nst = 0;
        __syncwarp();
    }
}

template <bool DIAG>
__device__ __forceinline__ unsigned scan_item(const Smem &S, const Params &P, int tid, int ioff0)
{
    double xi[VR], yi[VR], ti[VR];
    int ioff[VR];
#pragma unroll
    for (int r = 0; r < VR; ++r) {
        const double4 p = S.s_i[r * VT + tid];
        xi[r] = p.x;
        yi[r] = p.y;
        ti[r] = p.z;
        ioff[r] = ioff0 + r * VT;  // (global i) - (global j of sub-tile start)
    }
    const double tcut = P.tcut, tmax = P.tmax;
    unsigned npass = 0;
    int nst = 0;
    // 32-bit shared-window address of this thread's next free queue slot
    const unsigned qbase = (unsigned)__cvta_generic_to_shared(S.q + tid);
    unsigned qa = qbase;
    const unsigned qhigh = qbase + (QCAP - VG * VR) * VT * 2;
    for (int jb = 0; jb < TJ; jb += VG) {
        if (__any_sync(0xffffffffu, qa > qhigh)) {
            const int cnt = (int)((qa - qbase) / (2 * VT));
            drain(S, P, tid, cnt, nst, false);
            npass += cnt;
            qa = qbase;
        }
#pragma unroll
        for (int g = 0; g < VG; ++g) {
            const int j = jb + g;
            const double4 pj = S.s_j[j];
#pragma unroll
            for (int r = 0; r < VR; ++r) {
                const double dx = __dsub_rn(xi[r], pj.x);
                const double dy = __dsub_rn(yi[r], pj.y);
                const double d2 = __fma_rn(dy, dy, __dmul_rn(dx, dx));
                const double dt = fabs(__dsub_rn(ti[r], pj.z));
                bool pass = (d2 <= tcut) && (dt <= tmax);
                if (DIAG) pass = pass && (j > ioff[r]);
                if (pass) {
                    st_shared_u16(qa, (unsigned short)(j | (r << 8)));
                    qa += 2 * VT;
                }
            }
        }
    }
    const int cnt = (int)((qa - qbase) / (2 * VT));
    drain(S, P, tid, cnt, nst, true);
    npass += cnt;
    return npass;
}

__global__ void __launch_bounds__(VT, 3) pair_kernel(Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem S;
    S.s_i = reinterpret_cast<double4 *>(smem_raw);
    S.s_j = S.s_i + TI;
    S.q = reinterpret_cast<unsigned short *>(S.s_j + TJ);
    S.s_thr = reinterpret_cast<double *>(S.q + QCAP * VT);
    S.s_cnt = reinterpret_cast<unsigned *>(S.s_thr + P.n_space_bins + 2);
    S.stg = S.s_cnt + P.n_bins;

    const int tid = threadIdx.x;
    for (int b = tid; b < P.n_bins; b += VT) {
        S.s_cnt[b] = 0u;
    }
    for (int b = tid; b < P.n_space_bins + 2; b += VT) S.s_thr[b] = P.thr[b];
    unsigned long long *my_cnt = P.partial_cnt + (size_t)blockIdx.x * P.n_bins;
    double *my_sum = P.partial_sum + (size_t)blockIdx.x * P.n_bins;
    for (int b = tid; b < P.n_bins; b += VT) {
        my_cnt[b] = 0ull;
        my_sum[b] = 0.0;
    }
    S.g_sum = my_sum;

    const long long M = P.item_prefix[P.nI];  // total items
    unsigned long long npass = 0;
    unsigned items_done = 0, items_total = 0;
    // this partition owns items w with w % nparts == part; block b takes every grid-th
    for (long long k = blockIdx.x;; k += gridDim.x) {
        const long long w = k * P.nparts + P.part;
        if (w >= M) break;
        // item -> (I, J): largest I with prefix[I] <= w
        int lo = 0, hi = P.nI - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (P.item_prefix[mid] <= w) lo = mid; else hi = mid - 1;
        }
        const int I = lo;
        const int J = P.item_jbeg[I] + (int)(w - P.item_prefix[I]);

        __syncthreads();  // previous item finished with the tiles
        const double4 *gi = P.pts + (size_t)I * TI;
        const double4 *gj = P.pts + (size_t)J * TJ;
#pragma unroll
        for (int r = 0; r < VR; ++r) S.s_i[r * VT + tid] = gi[r * VT + tid];
        S.s_j[tid] = gj[tid];
        __syncthreads();

        const int ioff0 = (I * TI + tid) - J * TJ;
        if (J < I * JPI + JPI)
            npass += scan_item<true>(S, P, tid, ioff0);
        else
            npass += scan_item<false>(S, P, tid, ioff0);

        ++items_total;
        if (++items_done == FLUSH_ITEMS) {
            __syncthreads();
            for (int b = tid; b < P.n_bins; b += VT) {
                my_cnt[b] += S.s_cnt[b];
                S.s_cnt[b] = 0u;
            }
            items_done = 0;
        }
    }
    __syncthreads();
    for (int b = tid; b < P.n_bins; b += VT) {
        my_cnt[b] += S.s_cnt[b];
    }
    // stats: passing pairs (warp reduce, one atomic per warp)
    for (int off = 16; off > 0; off >>= 1) npass += __shfl_xor_sync(0xffffffffu, npass, off);
    if ((tid & 31) == 0 && npass) atomicAdd(&P.stats[2], npass);
    if (tid == 0 && items_total) atomicAdd(&P.stats[3], (unsigned long long)items_total);
}

// Fixed-order reduction over the per-block partials + the reference's peeled first pair.
__global__ void reduce_kernel(Params P, int nblocks, const double *x, const double *y,
                              const double *t, const double *v, int64_t n, double smax,
                              double ws, double wt, int peel, unsigned long long *counts,
                              double *sums)
{
    // 256 threads = 32 bins x 8 slices of the block list: slice s adds the partials of blocks
    // k = s, s+8, ... in increasing k (coalesced over the bins), then the 8 slices are combined
    // in a fixed tree -- the same order for any launch, so the result is deterministic.
    __shared__ double s_acc[8][32];
    __shared__ unsigned long long s_cnt8[8][32];
    const int bl = threadIdx.x & 31, ks = threadIdx.x >> 5;
    const int b = blockIdx.x * 32 + bl;
    {
        double a = 0.0;
        unsigned long long cc = 0;
        if (b < P.n_bins) {
#pragma unroll 4
            for (int k = ks; k < nblocks; k += 8) {
                a += P.partial_sum[(size_t)k * P.n_bins + b];
                cc += P.partial_cnt[(size_t)k * P.n_bins + b];
            }
        }
        s_acc[ks][bl] = a;
        s_cnt8[ks][bl] = cc;
    }
    __syncthreads();
    if (ks != 0 || b >= P.n_bins) return;
    double acc[8];
    unsigned long long c = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        acc[q] = s_acc[q][bl];
        c += s_cnt8[q][bl];
    }
    double s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
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
    counts[b] = c;
    sums[b] = s;
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

__global__ void __launch_bounds__(VT) sampled_kernel(Params P, SampledParams Q)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_sum = reinterpret_cast<double *>(smem_raw);
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_sum + P.n_bins);
    const int tid = threadIdx.x;
    for (int b = tid; b < P.n_bins; b += VT) {
        s_sum[b] = 0.0;
        s_cnt[b] = 0u;
    }
    unsigned long long *my_cnt = P.partial_cnt + (size_t)blockIdx.x * P.n_bins;
    double *my_sum = P.partial_sum + (size_t)blockIdx.x * P.n_bins;
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

    STIF_TRY(c->v_stats.reserve(4 * sizeof(unsigned long long)));
    STIF_CUDA(cudaMemsetAsync(c->v_stats.p, 0, 4 * sizeof(unsigned long long), st));
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

    // 1. sort indices by time
    STIF_TRY(c->v_idx.reserve(sizeof(unsigned) * n));
    STIF_TRY(c->v_idx_out.reserve(sizeof(unsigned) * n));
    const int tb = 256;
    const unsigned nb_n = (unsigned)((n + tb - 1) / tb);
    const unsigned *order = nullptr;
    if (cull) {
        STIF_TRY(c->v_keys.reserve(sizeof(double) * n));
        STIF_TRY(c->v_keys_out.reserve(sizeof(double) * n));
        pack_keys_kernel<<<nb_n, tb, 0, st>>>(t, n, c->v_keys.as<double>(), c->v_idx.as<unsigned>());
        size_t tmp_bytes = 0;
        STIF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, c->v_keys.as<double>(),
                                                  c->v_keys_out.as<double>(),
                                                  c->v_idx.as<unsigned>(),
                                                  c->v_idx_out.as<unsigned>(), (int)n, 0, 64, st));
        STIF_TRY(c->v_cubtmp.reserve(tmp_bytes));
        STIF_CUDA(cub::DeviceRadixSort::SortPairs(c->v_cubtmp.p, tmp_bytes, c->v_keys.as<double>(),
                                                  c->v_keys_out.as<double>(),
                                                  c->v_idx.as<unsigned>(),
                                                  c->v_idx_out.as<unsigned>(), (int)n, 0, 64, st));
        order = c->v_idx_out.as<unsigned>();
    } else {
        iota_kernel<<<nb_n, tb, 0, st>>>(n, c->v_idx.as<unsigned>());
        order = c->v_idx.as<unsigned>();
    }
    // 2. gather into sorted, padded AoS
    STIF_TRY(c->v_pts.reserve(sizeof(double4) * n_pad));
    gather_kernel<<<(unsigned)((n_pad + tb - 1) / tb), tb, 0, st>>>(x, y, t, v, order, n, n_pad,
                                                                    c->v_pts.as<double4>());
    // 3. work list
    STIF_TRY(c->v_tileinfo.reserve(sizeof(int) * (3 * (size_t)nI + 1)));
    int *count = c->v_tileinfo.as<int>();
    int *jbeg = count + nI;
    int *prefix = jbeg + nI;
    tile_range_kernel<<<(nI + 127) / 128, 128, 0, st>>>(c->v_pts.as<double4>(), n, nI, nJ, tmax,
                                                        cull ? 1 : 0, count, jbeg);
    scan_kernel<<<1, 1024, 0, st>>>(count, nI, prefix);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaEventRecord(c->ev[1], st));

    // 4. pair kernel
    Params P;
    P.pts = c->v_pts.as<double4>();
    P.item_prefix = prefix;
    P.item_jbeg = jbeg;
    P.nI = nI;
    P.n_bins = n_bins;
    P.n_time_bins = nt;
    P.n_space_bins = ns;
    P.mode = mode;
    P.part = part;
    P.nparts = nparts;
    P.tcut = sqrt_cut_threshold(smax);
    P.tmax = tmax;
    const double ws = smax / (double)ns, wt = tmax / (double)nt;  // utils.py:146-147
    P.inv_ws = 1.0 / ws;
    P.inv_wt = 1.0 / wt;
    P.mean = mean;
    P.stats = c->v_stats.as<unsigned long long>();
    P.inv_ws_f = (float)P.inv_ws;
    P.use_thr = (P.inv_ws > 0.0 && isfinite(P.inv_ws) && isfinite((double)P.inv_ws_f)) ? 1 : 0;
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

    const size_t smem = sizeof(double4) * (TI + TJ) + sizeof(unsigned short) * QCAP * VT +
                        sizeof(unsigned) * (size_t)n_bins +
                        sizeof(double) * ((size_t)ns + 2) + sizeof(unsigned) * 64 * (VT / 32);
    STIF_CUDA(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    int occ = 0;
    STIF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pair_kernel, VT, smem));
    if (occ < 1) {
        set_error("stif_variogram: pair kernel does not fit (smem %zu B)", smem);
        return STIF_ECUDA;
    }
    const int grid = c->sm_count * occ;
    STIF_TRY(c->v_partial.reserve((sizeof(double) + sizeof(unsigned long long)) * (size_t)grid *
                                  n_bins));
    P.partial_sum = c->v_partial.as<double>();
    P.partial_cnt = reinterpret_cast<unsigned long long *>(P.partial_sum + (size_t)grid * n_bins);
    pair_kernel<<<grid, VT, smem, st>>>(P);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaEventRecord(c->ev[2], st));

    // 5. reduce (+ peel fix-up on partition 0)
    const int peel = ((flags & STIF_VARIO_PEEL) && part == 0) ? 1 : 0;
    reduce_kernel<<<(n_bins + 31) / 32, 256, 0, st>>>(
        P, grid, x, y, t, v, n, smax, ws, wt, peel,
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
        STIF_CUDA(cudaMemcpyAsync(dx, x, in_bytes, cudaMemcpyHostToDevice, st));
        STIF_CUDA(cudaMemcpyAsync(dy, y, in_bytes, cudaMemcpyHostToDevice, st));
        STIF_CUDA(cudaMemcpyAsync(dt, t, in_bytes, cudaMemcpyHostToDevice, st));
        STIF_CUDA(cudaMemcpyAsync(dv, v, in_bytes, cudaMemcpyHostToDevice, st));
    }
    STIF_TRY(stif_variogram_dev(c, dx, dy, dt, dv, n, smax, tmax, ns, nt, mode, mean, flags, part,
                                nparts, dc, ds, st));
    STIF_CUDA(cudaMemcpyAsync(counts_host, dc, nb * 8, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaMemcpyAsync(sums_host, ds, nb * 8, cudaMemcpyDeviceToHost, st));
    STIF_CUDA(cudaStreamSynchronize(st));
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
    P.partial_sum = c->v_partial.as<double>();
    P.partial_cnt = reinterpret_cast<unsigned long long *>(P.partial_sum + (size_t)grid * n_bins);
    sampled_kernel<<<grid, VT, smem, st>>>(P, Q);
    STIF_CUDA(cudaGetLastError());
    STIF_CUDA(cudaEventRecord(c->ev[2], st));
    reduce_kernel<<<(n_bins + 31) / 32, 256, 0, st>>>(
        P, grid, x, y, t, v, n, smax, ws, wt, 0, reinterpret_cast<unsigned long long *>(counts_dev),
        sums_dev);
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
