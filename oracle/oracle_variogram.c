/*
 * oracle_variogram.c -- CPU restatement of the reference's empirical space-time
 * variogram / covariogram loop.  TEST INFRASTRUCTURE ONLY: imported by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker / the timed CPU baseline.  The product path never links or calls this file.
 *
 * Follows (reference file:line):
 *   stif/utils.py:87-101   _pair_index_generator  (all-pairs branch, i<j, i-major)
 *   stif/utils.py:104-190  get_variogram
 *   stif/utils.py:193-279  get_covariogram
 *
 * The reference is numba @njit(fastmath=True); what executes is LLVM -ffast-math
 * machine code, not IEEE-strict Python.  The arithmetic below is the EXECUTED form,
 * established by adversarial inputs against numba 0.65 (see oracle/gen_golden.py and
 * tests/golden/variogram_adversarial.npz):
 *   d2    = fma(dy, dy, rn(dx*dx))                 (FMA contraction)
 *   lag_s = sqrt_rn(d2)
 *   bin_s = trunc(lag_s * (1.0/ws)), bin_t = trunc(lag_t * (1.0/wt))
 *           (division by the loop-invariant width -> reciprocal multiply)
 *   EXCEPT the single peeled first iteration (i=0, j=1), which uses true division.
 *   hist[bin] += dv*dv is NOT contracted: rn(hist + rn(dv*dv)), sequential, i-major
 *   variogram = (hist*0.5)/norm
 * Every FP contraction is written explicitly; compile with -ffp-contract=off.
 *
 * Parity is pinned by tests/golden/*.npz, which hold outputs of the unmodified
 * reference run in the build container.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORACLE_MODE_VARIOGRAM 0
#define ORACLE_MODE_COVARIOGRAM 1

/* flags: bit0 = peel pair (0,1) with true division (the executed numba codegen; the
 *        reference is flags == ORACLE_FLAG_PEEL);
 *        bit1 = accumulate with fma(dv,dv,hist) (counter-factual)            */
#define ORACLE_FLAG_PEEL 1
#define ORACLE_FLAG_ACC_FMA 2
/* Counter-factual codegen variants, used ONLY by the fingerprint test to show that
 * the adversarial fixtures discriminate between them (the reference matches none). */
#define ORACLE_FLAG_D2_FMA_X 4   /* d2 = fma(dx,dx,rn(dy*dy)) */
#define ORACLE_FLAG_D2_UNFUSED 8 /* d2 = rn(dx*dx)+rn(dy*dy)  */
#define ORACLE_FLAG_TRUE_DIV 16  /* every pair uses lag/w      */

/* Raw accumulators (hist = sum of dv^2 or of (vi-mu)(vj-mu); norm = pair count as
 * double, like the reference).  x,y,t,v are SoA float64; int64 / bool inputs are
 * converted exactly by the caller.  Pairs i in [i_begin, i_end), j in (i, n). */
int64_t oracle_variogram_raw(const double *x, const double *y, const double *t,
                             const double *v, int64_t n, double space_dist_max,
                             double time_dist_max, int n_space_bins, int n_time_bins,
                             int mode, double mean, int flags, int64_t i_begin,
                             int64_t i_end, double *hist, double *norm)
{
    /* utils.py:146-147 */
    const double ws = space_dist_max / (double)n_space_bins;
    const double wt = time_dist_max / (double)n_time_bins;
    const double inv_ws = 1.0 / ws;
    const double inv_wt = 1.0 / wt;
    int64_t n_pairs = 0;
    if (i_end > n) i_end = n;
    for (int64_t i = i_begin; i < i_end; ++i) {
        const double xi = x[i], yi = y[i], ti = t[i], vi = v[i];
        for (int64_t j = i + 1; j < n; ++j) {
            ++n_pairs;
            /* utils.py:153-161 */
            const double dx = xi - x[j];
            const double dy = yi - y[j];
            double d2;
            if (flags & ORACLE_FLAG_D2_FMA_X)
                d2 = fma(dx, dx, dy * dy);
            else if (flags & ORACLE_FLAG_D2_UNFUSED)
                d2 = dx * dx + dy * dy;
            else
                d2 = fma(dy, dy, dx * dx);
            const double lag_s = sqrt(d2);
            if (lag_s > space_dist_max) continue;
            /* utils.py:162-164 */
            const double lag_t = fabs(ti - t[j]);
            if (lag_t > time_dist_max) continue;
            /* utils.py:165 / :240 */
            double a, b;
            if (mode == ORACLE_MODE_VARIOGRAM) {
                a = vi - v[j];
                b = a;
            } else {
                a = vi - mean;
                b = v[j] - mean;
            }
            /* utils.py:168-169 */
            double fs, ft;
            if ((flags & ORACLE_FLAG_TRUE_DIV) ||
                ((flags & ORACLE_FLAG_PEEL) && i == 0 && j == 1)) {
                fs = lag_s / ws;
                ft = lag_t / wt;
            } else {
                fs = lag_s * inv_ws;
                ft = lag_t * inv_wt;
            }
            const int64_t bs = (int64_t)fs;
            const int64_t bt = (int64_t)ft;
            /* utils.py:171-174 */
            if (bs >= 0 && bs < n_space_bins && bt >= 0 && bt < n_time_bins) {
                double *h = &hist[bs * n_time_bins + bt];
                if (flags & ORACLE_FLAG_ACC_FMA)
                    *h = fma(a, b, *h);
                else
                    *h = *h + a * b;
                norm[bs * n_time_bins + bt] += 1.0;
            }
        }
    }
    return n_pairs;
}

/* Full get_variogram / get_covariogram: returns variogram (NaN where norm==0),
 * norm, and the two bin widths (utils.py:178-190, :267-279).  `var` is only used
 * in covariogram mode (np.var(val), utils.py:235). */
void oracle_variogram(const double *x, const double *y, const double *t, const double *v,
                      int64_t n, double space_dist_max, double time_dist_max,
                      int n_space_bins, int n_time_bins, int mode, double mean, double var,
                      int flags, double *variogram, double *norm, double *widths)
{
    const int nb = n_space_bins * n_time_bins;
    memset(variogram, 0, sizeof(double) * (size_t)nb);
    memset(norm, 0, sizeof(double) * (size_t)nb);
    oracle_variogram_raw(x, y, t, v, n, space_dist_max, time_dist_max, n_space_bins,
                         n_time_bins, mode, mean, flags, 0, n, variogram, norm);
    for (int b = 0; b < nb; ++b) {
        if (norm[b] == 0.0) {
            variogram[b] = NAN;
        } else if (mode == ORACLE_MODE_VARIOGRAM) {
            /* executed form of hist/norm/2 under fastmath: (hist*0.5)/norm */
            variogram[b] = (variogram[b] * 0.5) / norm[b];
        } else {
            variogram[b] = (variogram[b] / norm[b]) / var;
        }
    }
    widths[0] = space_dist_max / (double)n_space_bins;
    widths[1] = time_dist_max / (double)n_time_bins;
}
