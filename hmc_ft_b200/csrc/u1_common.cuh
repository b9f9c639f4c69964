// Shared device/host helpers for the U(1) HMC kernels (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/u1hmc.h"
#include "u1_fastmath.cuh"

namespace u1 {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * kPi;

// Launch accounting (per host thread) so bench.py can report gpu_launches.
extern thread_local long long g_launches;
inline void count_launch(int n = 1) { g_launches += n; }

#define U1_CHECK_LAUNCH()                                  \
    do {                                                   \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)

#define U1_TRY(expr)                                       \
    do {                                                   \
        int s__ = (expr);                                  \
        if (s__ != 0) return s__;                          \
    } while (0)

// utils.py:182-187 regularize — the reference's exact expression (true division kept).
__device__ __forceinline__ double wrap_pi(double t) {
    double w = (t - kPi) / kTwoPi;
    return kTwoPi * (w - floor(w) - 0.5);
}

// Periodic index helpers for 0 <= i < L, shift in {-2..2}.
__device__ __forceinline__ int wrapi(int i, int L) {
    i = (i >= L) ? i - L : i;
    return (i < 0) ? i + L : i;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum of NV values per thread; result valid in every thread.
// `red` needs NV*32 doubles of shared memory.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();  // protect `red` against a previous use
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[k * 32 + wid] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = (lane < nw) ? red[k * 32 + lane] : 0.0;
        v[k] = warp_sum(t);
    }
}

// ---------------- Philox4x32-10 (Salmon et al. 2011) ----------------
struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
        U4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
        n.w = (uint32_t)p0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

__host__ __device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
    const uint64_t m = ((uint64_t)hi << 21) | ((uint64_t)lo >> 11);
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

struct PhiloxKey { uint32_t k0, k1; };
__host__ __device__ __forceinline__ PhiloxKey chain_key(uint64_t seed, uint64_t chain) {
    PhiloxKey k;
    k.k0 = (uint32_t)seed ^ (uint32_t)(chain * 0x9E3779B9ull);
    k.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)chain;
    return k;
}

// Table of u1::neg2_log (tools/gen_log_table.py): static constant data, one copy per translation unit.
static __device__ const LogTabEntry kLogTab[kLogTabN] = {
#include "u1_logtab.inc"
};

// Momenta of lattice site s (= x*L+y) for (seed, chain, traj): Box-Muller pair -> (pi0, pi1) =
// sqrt(-2 ln u1) (cos, sin)(2 pi u2).  Built from the branch-free helpers of u1_fastmath.cuh (about a
// third of the instructions of log()+sqrt()+sincos(); same values to a few 1e-16).  `tab` is kLogTab
// or a shared-memory copy of it.  The 1e-300 keeps u1 == 1 (probability 2^-53) finite.
__device__ __forceinline__ void philox_site_normals(PhiloxKey k, uint64_t traj, uint32_t site,
                                                    const LogTabEntry* __restrict__ tab, double& n0, double& n1) {
    U4 c{site, (uint32_t)traj, (uint32_t)(traj >> 32), 0u};
    U4 r = philox4x32_10(c, k.k0, k.k1);
    const double u1v = u01_53(r.x, r.y), u2v = u01_53(r.z, r.w);
    const double rad = sqrt_pos(neg2_log(u1v, tab) + 1e-300);
    double s, co;
    fast_sincos(kTwoPi * u2v, s, co);
    n0 = rad * co;
    n1 = rad * s;
}

__device__ __forceinline__ double philox_accept_uniform(PhiloxKey k, uint64_t traj) {
    U4 c{0xFFFFFFFFu, (uint32_t)traj, (uint32_t)(traj >> 32), 1u};
    U4 r = philox4x32_10(c, k.k0, k.k1);
    return u01_53(r.x, r.y);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Per-device caches (occupancy slots, one-time function attributes) are indexed by the CURRENT device, so a
// process that drives several GPUs gets one entry per GPU (include/u1hmc.h "Conventions").
constexpr int kMaxDevices = 64;
inline int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
    return d;
}
constexpr int kMaxGridY = 65535;     // chains ride in gridDim.y in several kernels: callers slice above this

}  // namespace u1
