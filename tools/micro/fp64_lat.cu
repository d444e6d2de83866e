// FP64 latency / issue microbenchmarks for sm_100a (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_lat fp64_lat.cu
// Prints cycles per dependent instruction for DFMA, DMUL->DFMA chains, MUFU.RSQ64H, DMMA.8x8x4,
// shared-memory round trips, and the DFMA throughput per SM as a function of warps per SM and
// independent chains per warp.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_dfma(double *out, long long *cyc, int iters, double a, double b)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = __fma_rn(x, a, b);
    }
    long long t1 = clock64();
    if (x == 1.2345) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lat_rsq(double *out, long long *cyc, int iters, double a)
{
    double x = 1.5 + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            double y;
            asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
            x = y + a;   // one DADD between MUFUs
        }
    }
    long long t1 = clock64();
    if (x == 1.2345) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lat_dmma(double *out, long long *cyc, int iters, double a, double b)
{
    double c0 = threadIdx.x, c1 = 1.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    if (c0 + c1 == 1.2345) out[0] = c0;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lat_lds(double *out, long long *cyc, int iters)
{
    __shared__ double s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (double)((i * 7 + 1) & 1023);
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) idx = (int)s[idx] ;   // LDS.64 -> F2I -> LDS
    }
    long long t1 = clock64();
    if (idx == -5) out[0] = idx;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// store -> load round trip through shared memory (same thread)
__global__ void lat_sts_lds(double *out, long long *cyc, int iters, double a)
{
    __shared__ double s[1024];
    double x = threadIdx.x;
    volatile double *vs = s;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            vs[threadIdx.x] = x;
            x = vs[threadIdx.x] + a;
        }
    }
    long long t1 = clock64();
    if (x == 1.2345) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lat_shfl(double *out, long long *cyc, int iters, double a)
{
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = __shfl_sync(0xffffffffu, x, (u * 5) & 31) + a;
    }
    long long t1 = clock64();
    if (x == 1.2345) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// throughput: CHAINS independent DFMA chains per thread
template <int CHAINS>
__global__ void thr_dfma(double *out, int iters, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = __fma_rn(x[c], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 1.2345) out[0] = s;
}

template <int CHAINS>
static void run_thr(double *d, int warps_per_sm, int sms)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    // one warp per block so that warps spread over the SM sub-partitions round-robin
    thr_dfma<CHAINS><<<sms * warps_per_sm, 32>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    thr_dfma<CHAINS><<<sms * warps_per_sm, 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)sms * warps_per_sm * iters * 8.0 * CHAINS;   // warp instructions
    double per_sm_per_clk = inst / sms / (ms * 1e-3 * 1.965e9);
    printf("  warps/SM %2d chains %d : %.3f warp-DFMA/clk/SM (peak 2.0)\n", warps_per_sm, CHAINS, per_sm_per_clk);
}

int main()
{
    double *d;
    long long *c, h;
    cudaMalloc(&d, 64);
    cudaMalloc(&c, 8);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int it = 2000;
#define LAT(name, ...)                                                                     \
    name<<<1, 32>>>(__VA_ARGS__);                                                          \
    name<<<1, 32>>>(__VA_ARGS__);                                                          \
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);                                          \
    printf("%-12s %.1f cycles per step\n", #name, (double)h / (it * 16.0));
    LAT(lat_dfma, d, c, it, 1.0000001, 1e-9)
    LAT(lat_rsq, d, c, it, 1.0)
    LAT(lat_dmma, d, c, it, 1e-3, 1e-3)
    LAT(lat_lds, d, c, it)
    LAT(lat_sts_lds, d, c, it, 1.0)
    LAT(lat_shfl, d, c, it, 1.0)
    printf("DFMA throughput (%d SMs):\n", p.multiProcessorCount);
    for (int w : {4, 8, 12, 16, 32}) {
        run_thr<1>(d, w, p.multiProcessorCount);
        run_thr<2>(d, w, p.multiProcessorCount);
        run_thr<4>(d, w, p.multiProcessorCount);
    }
    return 0;
}
