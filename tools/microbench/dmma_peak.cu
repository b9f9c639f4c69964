// FP64 tensor-core (DMMA m8n8k4) vs DFMA peak on B200.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    const double x = 1.0 - 1e-9, y = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], x, y);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 16 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks : {148 * 2, 148 * 4, 148 * 8}) {
        const int iters = 4096;
        k_dmma<<<blocks, 256>>>(out, 16);
        cudaEventRecord(e0); k_dmma<<<blocks, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double macs = (double)blocks * 8 /*warps*/ * iters * 8 * 256;
        printf("DMMA m8n8k4: blocks=%d  %.2f TFLOP/s (%s)\n", blocks, 2 * macs / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
        k_dfma<<<blocks, 256>>>(out, 16);
        cudaEventRecord(e0); k_dfma<<<blocks, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        macs = (double)blocks * 256 * iters * 32;
        printf("DFMA       : blocks=%d  %.2f TFLOP/s\n", blocks, 2 * macs / (ms * 1e-3) / 1e12);
    }
    return 0;
}
