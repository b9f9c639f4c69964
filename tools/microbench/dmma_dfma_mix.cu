// Do FP64 tensor-core MMAs (DMMA m8n8k4) and vector DFMAs share one pipe on B200, or can they overlap?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/dmma_dfma_mix tools/microbench/dmma_dfma_mix.cu
// Three kernels with the same launch shape (256 threads, 8 warps): DMMA only, DFMA only, and a mix in
// which every warp issues NM independent DMMAs and NF independent DFMAs per iteration.  If the pipes
// are separate the mixed kernel's combined TFLOP/s approaches the sum of the two peaks; if they are one
// pipe it stays at the single-pipe peak.  A second mix splits the work by warp (even warps DMMA, odd
// warps DFMA) to rule out an issue-slot limit inside one warp.
#include <cuda_runtime.h>
#include <cstdio>

template <int NM, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters) {
    double c[NM > 0 ? NM : 1][2], f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < NM; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    const double x = 1.0 - 1e-9, y = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NM > NF ? NM : NF); ++i) {
            if (i < NM)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            if (i < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(x), "d"(y));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

// warp-specialised mix: even warps DMMA only, odd warps DFMA only
template <int NM, int NF>
__global__ void __launch_bounds__(256) k_split(double* out, int iters_m, int iters_f) {
    double s = 0;
    if ((threadIdx.x >> 5) & 1) {
        double f[NF];
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = 1.0 + 1e-3 * (threadIdx.x + i);
        const double x = 1.0 - 1e-9, y = 1e-9;
        for (int it = 0; it < iters_f; ++it) {
#pragma unroll
            for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(x), "d"(y));
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) s += f[i];
    } else {
        double c[NM][2];
#pragma unroll
        for (int i = 0; i < NM; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
        const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
        for (int it = 0; it < iters_m; ++it) {
#pragma unroll
            for (int i = 0; i < NM; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
    }
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

static cudaEvent_t e0, e1;
template <typename F>
static float time_ms(F&& launch) {
    launch(16);
    cudaEventRecord(e0);
    launch(4096);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

template <int NM, int NF>
static void run_mix(double* out, int blocks) {
    const float ms = time_ms([&](int it) { k_mix<NM, NF><<<blocks, 256>>>(out, it); });
    const double warps = (double)blocks * 8, it = 4096;
    const double mma = 2.0 * warps * it * NM * 256, fma = 2.0 * warps * 32 * it * NF;
    printf("mix  NM=%2d NF=%2d blocks=%4d: %7.3f ms  DMMA %6.2f + DFMA %6.2f = %6.2f TFLOP/s (%s)\n", NM, NF, blocks, ms,
           mma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / 1e12, (mma + fma) / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
}

template <int NM, int NF>
static void run_split(double* out, int blocks, int ratio_f) {
    // iterations chosen so both halves take about the same time when the pipes are independent
    const float ms = time_ms([&](int it) { k_split<NM, NF><<<blocks, 256>>>(out, it, it * ratio_f); });
    const double warps = (double)blocks * 4, it = 4096;
    const double mma = 2.0 * warps * it * NM * 256, fma = 2.0 * warps * 32 * it * ratio_f * NF;
    printf("split NM=%2d NF=%2d x%d blocks=%4d: %7.3f ms  DMMA %6.2f + DFMA %6.2f = %6.2f TFLOP/s (%s)\n", NM, NF, ratio_f, blocks,
           ms, mma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / 1e12, (mma + fma) / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double* out;
    cudaMalloc(&out, 148 * 16 * 256 * 8);
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int blocks : {148 * 4, 148 * 8}) {
        run_mix<8, 0>(out, blocks);
        run_mix<0, 8>(out, blocks);
        run_mix<0, 16>(out, blocks);
        run_mix<8, 8>(out, blocks);      // flop ratio 8:1
        run_mix<4, 16>(out, blocks);     // 2:1
        run_mix<2, 16>(out, blocks);     // 1:1
        run_mix<1, 16>(out, blocks);     // 1:2
        run_split<8, 8>(out, blocks, 8);
        run_split<8, 8>(out, blocks, 4);
    }
    return 0;
}
