// Micro-benchmark: FP64 tensor-core mma.sync m8n8k4 throughput vs scalar DFMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a0, double b0)
{
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345) out[0] = s;
}
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b)
{
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = __fma_rn(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    if (s == 1.2345) out[0] = s;
}
int main()
{
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sm = 148, blocks = sm * 8, iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); dmma_kernel<<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = (double)blocks * 8 /*warps*/ * iters * 8.0 * (8 * 8 * 4 * 2);
        printf("DMMA m8n8k4: %.3f ms  %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
        cudaEventRecord(e0); dfma_kernel<<<blocks, 256>>>(d, iters / 8, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        flops = (double)blocks * 256 * (iters / 8) * 64.0 * 2;
        printf("DFMA scalar: %.3f ms  %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
