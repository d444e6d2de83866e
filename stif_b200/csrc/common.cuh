// common.cuh -- context, error handling and scratch management shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <vector>

#include "../../include/stif_b200.h"

namespace stif {

void set_error(const char *fmt, ...);

#define STIF_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            stif::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                \
                            cudaGetErrorString(e__));                                    \
            return (e__ == cudaErrorMemoryAllocation) ? STIF_ENOMEM : STIF_ECUDA;        \
        }                                                                                \
    } while (0)

#define STIF_TRY(call)                                                                   \
    do {                                                                                 \
        int r__ = (call);                                                                \
        if (r__ != STIF_OK) return r__;                                                  \
    } while (0)

// A grow-only device buffer.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return STIF_OK;
        if (p) {
            STIF_CUDA(cudaFree(p));
            p = nullptr;
            cap = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            cudaGetLastError();
            return STIF_ENOMEM;
        }
        cap = want;
        return STIF_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct KrigModel {
    int st_model = -1, space_model = 0, time_model = 0, metric_model = 0;
    int nparams = 0;
    double x[10] = {0};
};

}  // namespace stif

struct stif_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;      // the context's own stream
    cudaEvent_t ev[6] = {nullptr};      // phase timing events
    int n_phase = 0;                    // number of valid phases recorded in the last call
    cudaStream_t last_stream = nullptr; // stream the last call was enqueued on
    cudaStream_t stream2 = nullptr;     // copy stream of the chunked host pipeline
    cudaEvent_t v_upload_ev = nullptr;  // variogram host call: x, y, v uploaded (stream2)
    cudaEvent_t v_upload_xy_ev = nullptr;  // ... x and y uploaded (the space-time sort keys need them)
    bool v_wait_upload = false;         // stif_variogram_dev waits for it before the gather
    cudaEvent_t *ev_override = nullptr; // phase events of the chunk being enqueued (pipeline)
    bool keep_stats = false;            // pipeline: accumulate the statistics over the chunks
    std::vector<cudaEvent_t> chunk_ev;  // 4 phase events per chunk of the last pipelined call

    // ---- variogram scratch ----
    stif::DevBuf v_keys, v_keys_out, v_idx, v_idx_out, v_cubtmp;  // time sort
    stif::DevBuf v_pts;                                            // sorted AoS double4
    stif::DevBuf v_tileinfo;   // per i-tile j-range + prefix (work list)
    stif::DevBuf v_bb, v_items; // space-time sort mode: sub-tile boxes, explicit item list
    stif::DevBuf v_partial;    // per-block partial histograms
    stif::DevBuf v_stats;      // device counters
    stif::DevBuf v_stage;      // staging for host entry point
    stif::DevBuf v_thr;        // d2 bin thresholds
    std::vector<double> v_thr_host;
    uint64_t v_stats_host[4] = {0, 0, 0, 0};

    // ---- kriging state ----
    int64_t n_obs = 0;
    stif::DevBuf k_obs;        // obs SoA (x, y, t, resid) as given
    stif::DevBuf k_fit;        // fit objective: bins_space | bins_time | empirical | weights | out
    int fit_ns = 0, fit_nt = 0;
    double *fit_out_host = nullptr;   // pinned 8-byte landing buffer
    stif::DevBuf k_pairs_a, k_pairs_b, k_pairs_cnt;   // stand-alone neighbour pair list
    int64_t n_pairs = -1;
    stif::DevBuf k_pairtab;    // assembly pair table of the Cholesky kernel
    stif::DevBuf k_obs4;       // obs AoS {x, y, t, resid}, original order (neighbour gathers)
    stif::KrigModel model;
    stif::DevBuf k_scratch[16];
    stif::DevBuf k_stats;
    uint64_t k_stats_host[4] = {0, 0, 0, 0};
    unsigned long long *k_stats_pinned = nullptr;   // [8] statistics of the last host call (pinned)
    int k_chol_enabled = 0;    // last predict call tried the covariance-form kernel
    // grid cache key
    double grid_smax = -1, grid_tmax = -1;
    int64_t grid_nobs = -1;
};
