/*
 * oracle_kriging.c -- CPU restatement of the reference's space-time kriging path.
 * TEST INFRASTRUCTURE ONLY (checker in tests/ and smoke(); timed CPU baseline in
 * bench.py).  The product path never links or calls this file.
 *
 * Follows (reference file:line):
 *   stif/data.py:170-212        Data.get_kriging_idxs: dense filter over every
 *                               (observation, target) pair -- brute force, as the reference
 *   stif/predictor.py:602-660   nd_kriging: gamma-vector, keep the max_kriging_points
 *                               lowest-gamma candidates
 *   stif/predictor.py:566-593   calc_kriging_weights: bordered system, np.linalg.lstsq
 *   stif/variogram_models.py:7-218  spherical / gaussian x sum / product / product_sum /
 *                               metric / sum_metric
 *   stif/predictor.py:882-906   mean = sum(w32 * resid[idx]), std = sqrt(sum(w32 * g32))
 *
 * Third-party arithmetic the reference reaches on this path and how it is restated:
 *   - np.linalg.lstsq inside numba -> LAPACK dgelsd (rcond = -1 -> machine eps) from
 *     scipy's bundled OpenBLAS.  The SAME routine is called here through a function
 *     pointer supplied by the Python wrapper (scipy_dgelsd_ of libscipy_openblas), so the
 *     solver is not re-implemented.
 *   - scipy cdist 'sqeuclidean': unfused rn(dx*dx)+rn(dy*dy); 1-D 'euclidean' == |dt|.
 *   - numba argsort (unstable quicksort): ties at the cut are broken here by LOWEST
 *     OBSERVATION INDEX (north-star rule); `tie_out` flags targets where that matters.
 *   - numpy pairwise summation of the padded rows (np.sum(..., axis=1)).
 * The reference evaluates gamma(h,t) with LLVM fast-math; this file uses plain IEEE
 * double with the operation order of the Python source.  Agreement with the reference
 * is therefore ~1e-15 relative on gamma, not bitwise; the kriging tolerances are 1e-9.
 *
 * Parity is pinned by tests/golden/kriging_*.npz (outputs of the unmodified reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef void (*dgelsd_fn)(int *m, int *n, int *nrhs, double *a, int *lda, double *b, int *ldb,
                          double *s, double *rcond, int *rank, double *work, int *lwork,
                          int *iwork, int *info);

enum { ST_SUM = 0, ST_PRODUCT = 1, ST_PRODUCT_SUM = 2, ST_METRIC = 3, ST_SUM_METRIC = 4 };
enum { M_SPHERICAL = 0, M_GAUSSIAN = 1 };

typedef struct {
    int st, ks, kt, km;
    double x[10];
} oracle_model;

/* variogram_models.py:40-44 */
static double spherical(double h, double r, double c0, double b)
{
    const double a = r / 1.;
    if (h <= r) {
        const double u = h / a;
        return b + c0 * ((1.5 * u) - (0.5 * (u * u * u)));
    }
    return b + c0;
}

/* variogram_models.py:70-71 */
static double gaussian(double h, double r, double c0, double b)
{
    const double a = r / 2.;
    return b + c0 * (1. - exp(-((h * h) / (a * a))));
}

static double marginal(int kind, double h, double r, double c0, double b)
{
    return kind == M_SPHERICAL ? spherical(h, r, c0, b) : gaussian(h, r, c0, b);
}

double oracle_gamma(const oracle_model *M, double h, double t)
{
    const double *x = M->x;
    switch (M->st) {
    case ST_SUM: /* :98-99 */
        return marginal(M->ks, h, x[0], x[2], x[4]) + marginal(M->kt, t, x[1], x[3], x[5]);
    case ST_PRODUCT: /* :126-127 */
        return marginal(M->ks, h, x[0], x[2], x[4]) * marginal(M->kt, t, x[1], x[3], x[5]);
    case ST_PRODUCT_SUM: { /* :155-157 */
        const double p = marginal(M->ks, h, x[0], x[2], x[4]) * marginal(M->kt, t, x[1], x[3], x[5]);
        return x[6] * p + p;
    }
    case ST_METRIC: { /* :184-186 */
        const double d = sqrt(h * h + x[3] * x[3] * t * t);
        return marginal(M->km, d, x[0], x[1], x[2]);
    }
    default: { /* sum_metric :213-218 */
        const double d = sqrt(h * h + x[9] * x[9] * t * t);
        return (marginal(M->ks, h, x[0], x[3], x[6]) + marginal(M->kt, t, x[1], x[4], x[7])) +
               marginal(M->km, d, x[2], x[5], x[8]);
    }
    }
}

void oracle_gamma_vec(const oracle_model *M, const double *h, const double *t, int64_t n, double *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = oracle_gamma(M, h[i], t[i]);
}

/* numpy pairwise sum (numpy/_core/src/umath/loops_utils.h.src) + the reduction's initial 0 */
static double np_sum_f64(const double *a, int n)
{
    double res;
    if (n < 8) {
        res = 0.;
        for (int i = 0; i < n; ++i) res += a[i];
    } else if (n <= 128) {
        double r[8];
        int i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        res = np_sum_f64(a, n2) + np_sum_f64(a + n2, n - n2);
        return res; /* inner calls already added their 0 */
    }
    return 0. + res;
}

static float np_sum_f32(const float *a, int n)
{
    float res;
    if (n < 8) {
        res = 0.f;
        for (int i = 0; i < n; ++i) res += a[i];
    } else if (n <= 128) {
        float r[8];
        int i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        res = np_sum_f32(a, n2) + np_sum_f32(a + n2, n - n2);
        return res;
    }
    return 0.f + res;
}

typedef struct {
    double g;
    int64_t idx;
} cand_t;

static int cand_cmp(const void *pa, const void *pb)
{
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    if (a->g < b->g) return -1;
    if (a->g > b->g) return 1;
    return (a->idx > b->idx) - (a->idx < b->idx);
}

#define OK_F32_STORAGE 1

/*
 * Predict n_targets points.  Outputs: mean/std [n_targets]; optional (NULL to skip)
 * w_out/g_out float32 [n_targets*max_pts], idx_out uint32 [n_targets*max_pts] (zero padded),
 * n_used_out int32 (-1 = fewer than min_pts candidates), n_cand_out int32 (candidates in
 * the window), tie_out uint8 (1 = the max_pts-th and (max_pts+1)-th lowest gammas are equal,
 * i.e. the reference's unstable argsort may legitimately pick a different set),
 * rank_out int32 (dgelsd effective rank; < n_used+1 means rank-deficient -> min-norm).
 * Returns 0, or a LAPACK info code / -1 on allocation failure.
 */
int oracle_kriging(const double *ox, const double *oy, const double *ot, const double *resid,
                   int64_t n_obs, const double *tx, const double *ty, const double *tt,
                   const int64_t *leave_out, int64_t n_targets, int min_pts, int max_pts,
                   double space_dist_max, double time_dist_max, const oracle_model *M,
                   dgelsd_fn dgelsd, int flags, int n_threads, double *mean_out, double *std_out,
                   float *w_out, float *g_out, uint32_t *idx_out, int32_t *n_used_out,
                   int32_t *n_cand_out, uint8_t *tie_out, int32_t *rank_out)
{
    const double smax2 = space_dist_max * space_dist_max; /* data.py:207 */
    int status = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        int64_t cap = 1024;
        cand_t *cand = (cand_t *)malloc(sizeof(cand_t) * (size_t)cap);
        const int nmax = max_pts + 1;
        double *A = (double *)malloc(sizeof(double) * (size_t)nmax * nmax);
        double *b = (double *)malloc(sizeof(double) * (size_t)nmax);
        double *sv = (double *)malloc(sizeof(double) * (size_t)nmax);
        double *t64 = (double *)malloc(sizeof(double) * (size_t)max_pts);
        float *w32 = (float *)malloc(sizeof(float) * (size_t)max_pts);
        float *g32 = (float *)malloc(sizeof(float) * (size_t)max_pts);
        float *t32 = (float *)malloc(sizeof(float) * (size_t)max_pts);
        /* workspace query at the largest size */
        int lwork = -1, liwork = 0, info = 0, rank = 0, one = 1, nn = nmax;
        double wq = 0, rcond = -1.0;
        int iwq = 0;
        dgelsd(&nn, &nn, &one, A, &nn, b, &nn, sv, &rcond, &rank, &wq, &lwork, &iwq, &info);
        lwork = (int)wq + 64;
        liwork = iwq + 64;
        if (liwork < 3 * nmax * 12 + 11 * nmax + 64) liwork = 3 * nmax * 12 + 11 * nmax + 64;
        double *work = (double *)malloc(sizeof(double) * (size_t)lwork);
        int *iwork = (int *)malloc(sizeof(int) * (size_t)liwork);
        if (!cand || !A || !b || !sv || !t64 || !w32 || !g32 || !t32 || !work || !iwork) {
#pragma omp critical
            status = -1;
        }

#pragma omp for schedule(dynamic, 16)
        for (int64_t ti = 0; ti < n_targets; ++ti) {
            if (status) continue;
            const double x = tx[ti], y = ty[ti], t = tt[ti];
            const int64_t leave = leave_out ? leave_out[ti] : -1;
            /* data.py:195-212: dense filter over all observations, ascending index */
            int64_t m = 0;
            for (int64_t o = 0; o < n_obs; ++o) {
                const double dx = ox[o] - x, dy = oy[o] - y;
                const double d2 = dx * dx + dy * dy;
                if (!(d2 < smax2)) continue;
                const double dt = fabs(ot[o] - t);
                if (!(dt < time_dist_max)) continue;
                if (!(ot[o] <= t)) continue;
                if (o == leave) continue;
                if (m == cap) {
                    cap *= 2;
                    cand_t *nc = (cand_t *)realloc(cand, sizeof(cand_t) * (size_t)cap);
                    if (!nc) {
#pragma omp critical
                        status = -1;
                        break;
                    }
                    cand = nc;
                }
                cand[m].idx = o;
                cand[m].g = 0.0;
                ++m;
            }
            if (status) continue;
            if (n_cand_out) n_cand_out[ti] = (int32_t)m;
            float *wrow = w_out ? w_out + (size_t)ti * max_pts : NULL;
            float *grow = g_out ? g_out + (size_t)ti * max_pts : NULL;
            uint32_t *irow = idx_out ? idx_out + (size_t)ti * max_pts : NULL;
            if (wrow) memset(wrow, 0, sizeof(float) * (size_t)max_pts);
            if (grow) memset(grow, 0, sizeof(float) * (size_t)max_pts);
            if (irow) memset(irow, 0, sizeof(uint32_t) * (size_t)max_pts);
            if (tie_out) tie_out[ti] = 0;
            if (rank_out) rank_out[ti] = 0;
            /* predictor.py:624-626 */
            if (m < min_pts || m == 0) {
                mean_out[ti] = 0.0;
                std_out[ti] = 0.0;
                if (n_used_out) n_used_out[ti] = m < min_pts ? -1 : 0;
                continue;
            }
            /* predictor.py:627-636 */
            for (int64_t e = 0; e < m; ++e) {
                const int64_t o = cand[e].idx;
                const double dx = ox[o] - x, dy = oy[o] - y;
                const double h = sqrt(dx * dx + dy * dy);
                const double lt = fabs(ot[o] - t);
                cand[e].g = oracle_gamma(M, h, lt);
            }
            /* predictor.py:637-642 (ties: lowest index) */
            if (m > max_pts) {
                qsort(cand, (size_t)m, sizeof(cand_t), cand_cmp);
                if (tie_out && cand[max_pts - 1].g == cand[max_pts].g) tie_out[ti] = 1;
                m = max_pts;
            }
            const int k = (int)m, n = k + 1;
            /* predictor.py:573-585, column-major for LAPACK (A is symmetric) */
            for (int a = 0; a < k; ++a) {
                const int64_t oa = cand[a].idx;
                for (int c = 0; c < k; ++c) {
                    const int64_t oc = cand[c].idx;
                    const double dx = ox[oa] - ox[oc], dy = oy[oa] - oy[oc];
                    const double h = sqrt(dx * dx + dy * dy); /* utils.py:40-44 */
                    const double lt = fabs(ot[oa] - ot[oc]);   /* utils.py:19-22 */
                    A[(size_t)c * n + a] = oracle_gamma(M, h, lt);
                }
                A[(size_t)k * n + a] = 1.0;
                A[(size_t)a * n + k] = 1.0;
                b[a] = cand[a].g;
            }
            A[(size_t)k * n + k] = 0.0;
            b[k] = 1.0;
            /* predictor.py:587: np.linalg.lstsq -> dgelsd, rcond=-1 */
            int nn2 = n, lw = lwork;
            rcond = -1.0;
            dgelsd(&nn2, &nn2, &one, A, &nn2, b, &nn2, sv, &rcond, &rank, work, &lw, iwork, &info);
            if (info != 0) {
#pragma omp critical
                status = info;
                continue;
            }
            if (rank_out) rank_out[ti] = rank;
            if (n_used_out) n_used_out[ti] = k;
            /* predictor.py:646-658 + :900-905 */
            if (flags & OK_F32_STORAGE) {
                for (int e = 0; e < max_pts; ++e) {
                    float w = 0.f, g = 0.f;
                    double r = resid[0];
                    uint32_t id = 0;
                    if (e < k) {
                        w = (float)b[e];
                        g = (float)cand[e].g;
                        id = (uint32_t)cand[e].idx;
                        r = resid[id];
                    }
                    w32[e] = w;
                    g32[e] = g;
                    t64[e] = (double)w * r;
                    t32[e] = w * g;
                    if (wrow) wrow[e] = w;
                    if (grow) grow[e] = g;
                    if (irow) irow[e] = id;
                }
                mean_out[ti] = np_sum_f64(t64, max_pts);
                std_out[ti] = (double)sqrtf(np_sum_f32(t32, max_pts));
            } else {
                double am = 0.0, av = 0.0;
                for (int e = 0; e < k; ++e) {
                    am += b[e] * resid[cand[e].idx];
                    av += b[e] * cand[e].g;
                    if (wrow) wrow[e] = (float)b[e];
                    if (grow) grow[e] = (float)cand[e].g;
                    if (irow) irow[e] = (uint32_t)cand[e].idx;
                }
                mean_out[ti] = am;
                std_out[ti] = sqrt(av);
            }
        }
        free(cand);
        free(A);
        free(b);
        free(sv);
        free(t64);
        free(w32);
        free(g32);
        free(t32);
        free(work);
        free(iwork);
    }
    return status;
}

/* Float64 weights of one system (for solver-level tests): A is built from the given
 * neighbour coordinates exactly as calc_kriging_weights does. */
int oracle_kriging_weights(const double *xs, const double *ys, const double *ts,
                           const double *gvec, int k, const oracle_model *M, dgelsd_fn dgelsd,
                           double *w_out, int *rank_out)
{
    const int n = k + 1;
    double *A = (double *)malloc(sizeof(double) * (size_t)n * n);
    double *b = (double *)malloc(sizeof(double) * (size_t)n);
    double *sv = (double *)malloc(sizeof(double) * (size_t)n);
    int lwork = -1, info = 0, rank = 0, one = 1, nn = n, iwq = 0;
    double wq = 0, rcond = -1.0;
    dgelsd(&nn, &nn, &one, A, &nn, b, &nn, sv, &rcond, &rank, &wq, &lwork, &iwq, &info);
    lwork = (int)wq + 64;
    int liwork = iwq + 3 * n * 12 + 11 * n + 64;
    double *work = (double *)malloc(sizeof(double) * (size_t)lwork);
    int *iwork = (int *)malloc(sizeof(int) * (size_t)liwork);
    if (!A || !b || !sv || !work || !iwork) return -1;
    for (int a = 0; a < k; ++a) {
        for (int c = 0; c < k; ++c) {
            const double dx = xs[a] - xs[c], dy = ys[a] - ys[c];
            A[(size_t)c * n + a] = oracle_gamma(M, sqrt(dx * dx + dy * dy), fabs(ts[a] - ts[c]));
        }
        A[(size_t)k * n + a] = 1.0;
        A[(size_t)a * n + k] = 1.0;
        b[a] = gvec[a];
    }
    A[(size_t)k * n + k] = 0.0;
    b[k] = 1.0;
    dgelsd(&nn, &nn, &one, A, &nn, b, &nn, sv, &rcond, &rank, work, &lwork, iwork, &info);
    for (int a = 0; a < k; ++a) w_out[a] = b[a];
    if (rank_out) *rank_out = rank;
    free(A);
    free(b);
    free(sv);
    free(work);
    free(iwork);
    return info;
}
