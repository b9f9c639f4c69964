// Domain-decomposed plain HMC (BASELINE configs[4]): halo exchange and the cross-rank Metropolis decision
// done by OUR kernels over NVLink peer memory - no NCCL call, no host round trip on the data path, so the
// whole trajectory of a slab (about 30 launches) is one CUDA graph on every rank.
//
// Every rank owns a small "peer region" (cudaMalloc + CUDA IPC, opened by the other ranks of the box; NVSwitch
// gives every GPU load/store access to every peer): flags and counters, an all-gather slot for the
// Hamiltonian sums of every rank, and a double-buffered staging area for incoming halo rows.
//
//   k_slab_exchange        push my edge rows (theta, optionally pi) straight into the neighbours' staging
//                          slots with 16-byte stores over NVLink, publish an epoch flag with st.release.sys,
//                          wait for the neighbours' flags (ld.acquire.sys on LOCAL memory) and unpack the
//                          staged rows into my halo rows - one launch per exchange, all CTAs co-resident.
//   k_slab_metropolis_peer every rank writes its 6 Hamiltonian sums into every peer's slot, waits for all
//                          slots, adds them in rank order (so all ranks form bit-identical totals) and takes
//                          the accept/reject decision from the shared Philox draw (hmc_u1.py:84-97).
//
// Epochs live in device counters that the kernels advance themselves, so a captured graph replays correctly.
// Staging and sum slots are double-buffered by epoch parity: a slot written at exchange e was last read at
// exchange e-2, and the writer can only have reached e after the reader finished e-1 (flag dependency chain).
// All waits are bounded (kPeerTimeoutNs): a dead peer raises the region's error word instead of hanging the GPU.
#include "u1_common.cuh"
#include <cstring>

namespace u1 {

// ---- peer region layout (u64 words at the start, then sums, then staging) ----
enum : int { W_F_TOP = 0, W_F_BOT = 1, W_C_XCH = 2, W_C_PUSH = 3, W_C_DONE = 4, W_C_MTR = 5, W_ERR = 6, W_F_SUM0 = 16 };
constexpr int kPeerMaxWorld = 16;
constexpr size_t kPeerSumsOff = 512;          // [2 parity][16 ranks][8 doubles]
constexpr size_t kPeerStageOff = 4096;        // [2 parity][2 sides: 0 top-in, 1 bottom-in][4 planes][k*L doubles]
constexpr unsigned long long kPeerTimeoutNs = 4000000000ull;   // per wait; the error word is sticky (later waits are skipped)

__host__ __device__ inline size_t peer_stage_doubles(int k, int L) { return (size_t)4 * k * L; }   // one side, one parity

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= want (flag in LOCAL memory, written by a peer over NVLink); false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long want) {
    if (ld_acquire_sys(flag) >= want) return true;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < want) {
        __nanosleep(100);
        if (global_ns() - t0 > kPeerTimeoutNs) return false;
    }
    return true;
}

struct XchArgs {
    double* ext_theta;          // [2][Lx+2k][L]
    double* ext_pi;             // same layout, or nullptr (theta-only exchange)
    unsigned long long* self;   // my region
    unsigned long long* up;     // region of rank-1 (receives my FIRST rows as its bottom halo)
    unsigned long long* dn;     // region of rank+1 (receives my LAST rows as its top halo)
    int Lx, k, L, rows_up, rows_dn;
};

constexpr int kXchThreads = 256;
constexpr int kXchUnroll = 4;

__global__ void __launch_bounds__(kXchThreads) k_slab_exchange(XchArgs a) {
    __shared__ unsigned long long s_epoch;
    __shared__ int s_ok;
    const int tid = threadIdx.x;
    if (tid == 0) s_epoch = *reinterpret_cast<volatile unsigned long long*>(a.self + W_C_XCH);
    __syncthreads();
    const unsigned long long e = s_epoch;
    const int par = (int)(e & 1ull);
    const int L = a.L, k = a.k, Lx = a.Lx;
    const int nplanes = a.ext_pi ? 4 : 2;
    const size_t plane = (size_t)(Lx + 2 * k) * L;                 // doubles per mu plane of an ext buffer
    const size_t side = peer_stage_doubles(k, L);
    const int L2 = L >> 1;                                          // double2 units per row
    auto src_plane = [&](int p) -> double* { return (p < 2 ? a.ext_theta : a.ext_pi) + (size_t)(p & 1) * plane; };
    double* st_up = reinterpret_cast<double*>(reinterpret_cast<char*>(a.up) + kPeerStageOff) + ((size_t)par * 2 + 1) * side;   // up's bottom-in
    double* st_dn = reinterpret_cast<double*>(reinterpret_cast<char*>(a.dn) + kPeerStageOff) + ((size_t)par * 2 + 0) * side;   // dn's top-in
    // ---- push: my first rows_up rows -> up ; my last rows_dn rows -> dn ----
    // (kXchUnroll independent 16-byte loads in flight per thread before the first store: the copy is latency-bound)
    const long long n_up = (long long)nplanes * a.rows_up * L2, n_dn = (long long)nplanes * a.rows_dn * L2;
    const long long stride = (long long)gridDim.x * kXchThreads;
    for (long long i0 = blockIdx.x * (long long)kXchThreads + tid; i0 < n_up + n_dn; i0 += stride * kXchUnroll) {
        double2 v[kXchUnroll];
        double* dst[kXchUnroll];
#pragma unroll
        for (int u = 0; u < kXchUnroll; ++u) {
            const long long i = i0 + u * stride;
            dst[u] = nullptr;
            if (i < n_up + n_dn) {
                const bool to_up = i < n_up;
                const long long j = to_up ? i : i - n_up;
                const int rows = to_up ? a.rows_up : a.rows_dn;
                const int p = (int)(j / ((long long)rows * L2));
                const int rem = (int)(j - (long long)p * rows * L2);
                const int r = rem / L2, c = (rem - r * L2) << 1;
                const int row = to_up ? (k + r) : (Lx + k - rows + r);
                v[u] = *reinterpret_cast<const double2*>(src_plane(p) + (size_t)row * L + c);
                dst[u] = (to_up ? st_up : st_dn) + ((size_t)p * k + r) * L + c;
            }
        }
#pragma unroll
        for (int u = 0; u < kXchUnroll; ++u)
            if (dst[u]) *reinterpret_cast<double2*>(dst[u]) = v[u];             // 16-byte store over NVLink
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        const unsigned long long old = atomicAdd(a.self + W_C_PUSH, 1ull);
        if (old == gridDim.x - 1) {                                 // last CTA: every push of this rank is visible
            __threadfence_system();
            a.self[W_C_PUSH] = 0ull;
            st_release_sys(a.up + W_F_BOT, e + 1);
            st_release_sys(a.dn + W_F_TOP, e + 1);
        }
        // ---- wait for the neighbours' rows ----
        bool ok = *reinterpret_cast<volatile unsigned long long*>(a.self + W_ERR) == 0ull;     // sticky: no more waiting after a timeout
        ok = ok && wait_flag(a.self + W_F_TOP, e + 1);
        ok = ok && wait_flag(a.self + W_F_BOT, e + 1);
        if (!ok) atomicCAS(a.self + W_ERR, 0ull, 1ull);          // keep the FIRST error code
        s_ok = ok ? 1 : 0;
    }
    __syncthreads();
    // ---- unpack: top halo <- up's last rows_dn rows ; bottom halo <- dn's first rows_up rows ----
    const double* in_top = reinterpret_cast<const double*>(reinterpret_cast<const char*>(a.self) + kPeerStageOff) + ((size_t)par * 2 + 0) * side;
    const double* in_bot = in_top + side;
    const long long m_top = (long long)nplanes * a.rows_dn * L2, m_bot = (long long)nplanes * a.rows_up * L2;
    if (s_ok) {
        for (long long i0 = blockIdx.x * (long long)kXchThreads + tid; i0 < m_top + m_bot; i0 += stride * kXchUnroll) {
            double2 v[kXchUnroll];
            double* dst[kXchUnroll];
#pragma unroll
            for (int u = 0; u < kXchUnroll; ++u) {
                const long long i = i0 + u * stride;
                dst[u] = nullptr;
                if (i < m_top + m_bot) {
                    const bool top = i < m_top;
                    const long long j = top ? i : i - m_top;
                    const int rows = top ? a.rows_dn : a.rows_up;
                    const int p = (int)(j / ((long long)rows * L2));
                    const int rem = (int)(j - (long long)p * rows * L2);
                    const int r = rem / L2, c = (rem - r * L2) << 1;
                    const int row = top ? (k - rows + r) : (Lx + k + r);
                    v[u] = *reinterpret_cast<const double2*>((top ? in_top : in_bot) + ((size_t)p * k + r) * L + c);
                    dst[u] = src_plane(p) + (size_t)row * L + c;
                }
            }
#pragma unroll
            for (int u = 0; u < kXchUnroll; ++u)
                if (dst[u]) *reinterpret_cast<double2*>(dst[u]) = v[u];
        }
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned long long old = atomicAdd(a.self + W_C_DONE, 1ull);
        if (old == gridDim.x - 1) {                                 // last CTA: advance the epoch for the next exchange
            a.self[W_C_DONE] = 0ull;
            __threadfence();
            *reinterpret_cast<volatile unsigned long long*>(a.self + W_C_XCH) = e + 1;
        }
    }
}

struct PeerTable { unsigned long long* region[kPeerMaxWorld]; };

// One CTA.  sums_old / sums_new: this rank's {sum cos P, sum wrap P, sum pi^2}.  Outputs as u1_metropolis_sums.
__global__ void __launch_bounds__(32)
k_slab_metropolis_peer(PeerTable pt, int rank, int world, const double* __restrict__ s_old, const double* __restrict__ s_new,
                       const double* __restrict__ u_inject, uint64_t seed, uint64_t chain, uint64_t traj,
                       const uint64_t* __restrict__ traj_dev, double beta, double inv_vol,
                       double* out_dH, int32_t* out_acc, double* out_H, double* out_plaq, double* out_topo) {
    const int r = threadIdx.x;
    unsigned long long* self = pt.region[rank];
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(self + W_C_MTR);
    const int par = (int)(e & 1ull);
    if (r < world) {                                    // publish my sums in peer r's slot [par][rank]
        double* slot = reinterpret_cast<double*>(reinterpret_cast<char*>(pt.region[r]) + kPeerSumsOff) + ((size_t)par * kPeerMaxWorld + rank) * 8;
#pragma unroll
        for (int i = 0; i < 3; ++i) { slot[i] = s_old[i]; slot[3 + i] = s_new[i]; }
        __threadfence_system();
        st_release_sys(pt.region[r] + W_F_SUM0 + rank, e + 1);
        if (*reinterpret_cast<volatile unsigned long long*>(self + W_ERR) == 0ull && !wait_flag(self + W_F_SUM0 + r, e + 1))
            atomicCAS(self + W_ERR, 0ull, 2ull);
    }
    __syncwarp();
    if (r == 0) {
        __threadfence_system();
        const double* in = reinterpret_cast<const double*>(reinterpret_cast<const char*>(self) + kPeerSumsOff) + (size_t)par * kPeerMaxWorld * 8;
        double o[3] = {0.0, 0.0, 0.0}, n[3] = {0.0, 0.0, 0.0};
        for (int q = 0; q < world; ++q)                 // fixed rank order: identical totals on every rank
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                o[i] += *reinterpret_cast<const volatile double*>(in + q * 8 + i);
                n[i] += *reinterpret_cast<const volatile double*>(in + q * 8 + 3 + i);
            }
        if (traj_dev) traj += *traj_dev;
        const double H_old = (-beta) * o[0] + 0.5 * o[2];
        const double H_new = (-beta) * n[0] + 0.5 * n[2];
        const double dH = H_new - H_old;
        const double u = u_inject ? u_inject[0] : philox_accept_uniform(chain_key(seed, chain), traj);
        const bool acc = u < exp(-dH);
        const double* f = acc ? n : o;
        if (out_dH) out_dH[0] = dH;
        out_acc[0] = acc ? 1 : 0;
        if (out_H) out_H[0] = acc ? H_new : H_old;
        if (out_plaq) out_plaq[0] = f[0] * inv_vol;
        if (out_topo) out_topo[0] = floor(0.1 + f[1] / kTwoPi);
        *reinterpret_cast<volatile unsigned long long*>(self + W_C_MTR) = e + 1;
    }
}

// out[mu][i] = wrap(theta[mu][i] + a*pi[mu][i]) for both mu planes of a halo-extended buffer in one launch
__global__ void k_drift_wrap_planes(double* th, const double* __restrict__ pi, size_t plane, size_t n, double a) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 2 * n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t mu = i / n, j = i - mu * n;
        th[mu * plane + j] = wrap_pi(fma(a, pi[mu * plane + j], th[mu * plane + j]));   // exact utils.py:182-187 formula
    }
}

// theta_own[mu][i] = accept ? ext[mu][i] : theta_own[mu][i]   /   ext[mu][i] = theta_own[mu][i]  (copy-in)
__global__ void k_slab_select_planes(double* __restrict__ own, const double* __restrict__ ext, size_t plane, size_t n,
                                     const int32_t* __restrict__ acc) {
    if (!acc[0]) return;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 2 * n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t mu = i / n, j = i - mu * n;
        own[i] = ext[mu * plane + j];
    }
}
// on accept: state[mu][j] = src[mu][j] (both plane-strided) and, if given, the caller's contiguous copy own[i] too
// sums_state[0..1] (sum cos P, sum wrap P of the state: the action part of the next H_old) follow the state
__global__ void k_slab_select_state(double* __restrict__ state, const double* __restrict__ src, size_t plane, size_t n,
                                    double* __restrict__ own, const int32_t* __restrict__ acc,
                                    double* __restrict__ sums_state, const double* __restrict__ sums_new) {
    if (!acc[0]) return;
    if (sums_state && blockIdx.x == 0 && threadIdx.x == 0) { sums_state[0] = sums_new[0]; sums_state[1] = sums_new[1]; }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 2 * n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t mu = i / n, j = i - mu * n;
        const double v = src[mu * plane + j];
        state[mu * plane + j] = v;
        if (own) own[i] = v;
    }
}
__global__ void k_slab_copy_in(const double* __restrict__ own, double* __restrict__ ext, size_t plane, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 2 * n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t mu = i / n, j = i - mu * n;
        ext[mu * plane + j] = own[i];
    }
}

static inline int slab_grid(size_t n) {
    size_t g = (n + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace u1

using namespace u1;

extern "C" {

size_t u1_ipc_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

size_t u1_slab_peer_region_bytes(int k, int L) {
    if (k <= 0 || L <= 0) return 0;
    return kPeerStageOff + (size_t)2 * 2 * peer_stage_doubles(k, L) * sizeof(double);
}

int u1_peer_alloc(void** dev_ptr, size_t bytes, void* ipc_handle_out) {
    if (!dev_ptr || bytes == 0) return U1_ERR_ARG;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess && ipc_handle_out) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle_out), p);
    if (e != cudaSuccess) { cudaFree(p); return (int)e; }
    *dev_ptr = p;
    return U1_OK;
}

int u1_peer_open(const void* ipc_handle, void** dev_ptr) {
    if (!ipc_handle || !dev_ptr) return U1_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? U1_OK : (int)e;
}

int u1_peer_close(void* dev_ptr) {
    if (!dev_ptr) return U1_ERR_ARG;
    cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    return e == cudaSuccess ? U1_OK : (int)e;
}

int u1_peer_free(void* dev_ptr) {
    if (!dev_ptr) return U1_ERR_ARG;
    cudaError_t e = cudaFree(dev_ptr);
    return e == cudaSuccess ? U1_OK : (int)e;
}

// With lazy module loading (the CUDA 12 default) the FIRST launch of a kernel loads its code, and that load can
// wait for the device to go idle.  A kernel that spins on a peer's flag must therefore never be the reason a
// not-yet-loaded kernel of the same trajectory cannot start: load everything the slab trajectory uses up front.
int u1_plain_preload_slab_kernels(void);                    // u1_plain.cu
int u1_slab_preload(void) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, k_slab_exchange);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_slab_metropolis_peer);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_drift_wrap_planes);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_slab_select_planes);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_slab_copy_in);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_slab_select_state);
    if (e != cudaSuccess) return (int)e;
    return u1_plain_preload_slab_kernels();
}

int u1_slab_peer_status(const void* region_self, unsigned long long* host_words8) {
    if (!region_self || !host_words8) return U1_ERR_ARG;
    cudaError_t e = cudaMemcpy(host_words8, region_self, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? U1_OK : (int)e;
}

int u1_slab_exchange(double* ext_theta, double* ext_pi, int Lx, int k, int L, int rows_up, int rows_dn,
                     void* region_self, void* region_up, void* region_dn, u1_stream_t stream) {
    g_launches = 0;
    if (!ext_theta || !region_self || !region_up || !region_dn || Lx <= 0 || k <= 0 || L <= 0) return U1_ERR_ARG;
    if (rows_up < 0 || rows_dn < 0 || rows_up > k || rows_dn > k || k > Lx || (L & 1)) return U1_ERR_ARG;
    if (!aligned16(ext_theta) || (ext_pi && !aligned16(ext_pi))) return U1_ERR_ARG;
    XchArgs a{};
    a.ext_theta = ext_theta; a.ext_pi = ext_pi;
    a.self = (unsigned long long*)region_self; a.up = (unsigned long long*)region_up; a.dn = (unsigned long long*)region_dn;
    a.Lx = Lx; a.k = k; a.L = L; a.rows_up = rows_up; a.rows_dn = rows_dn;
    const long long units = (long long)(ext_pi ? 4 : 2) * (rows_up + rows_dn) * (L >> 1);
    long long g = (units + kXchThreads * 4 - 1) / (kXchThreads * 4);          // ~4 double2 per thread
    if (g < 1) g = 1;
    if (g > 120) g = 120;                                                        // all CTAs co-resident (spin waits inside)
    k_slab_exchange<<<(int)g, kXchThreads, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_slab_metropolis_peer(const double* sums_old, const double* sums_new, const double* u_inject,
                            uint64_t seed, uint64_t chain, uint64_t traj, const uint64_t* traj_dev,
                            double beta, double volume, int rank, int world, void* const* regions,
                            double* out_dH, int32_t* out_accept, double* out_H, double* out_plaq, double* out_topo,
                            u1_stream_t stream) {
    g_launches = 0;
    if (!sums_old || !sums_new || !out_accept || !regions || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world ||
        !(volume > 0))
        return U1_ERR_ARG;
    PeerTable pt{};
    for (int r = 0; r < world; ++r) {
        if (!regions[r]) return U1_ERR_ARG;
        pt.region[r] = (unsigned long long*)regions[r];
    }
    k_slab_metropolis_peer<<<1, 32, 0, (cudaStream_t)stream>>>(pt, rank, world, sums_old, sums_new, u_inject, seed, chain, traj,
                                                                traj_dev, beta, 1.0 / volume, out_dH, out_accept, out_H,
                                                                out_plaq, out_topo);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_slab_drift_wrap(double* ext_theta_own, const double* ext_pi_own, size_t plane_stride, int Lx, int L, double a,
                       u1_stream_t stream) {
    g_launches = 0;
    if (!ext_theta_own || !ext_pi_own || Lx <= 0 || L <= 0 || plane_stride < (size_t)Lx * L) return U1_ERR_ARG;
    const size_t n = (size_t)Lx * L;
    k_drift_wrap_planes<<<slab_grid(2 * n), 256, 0, (cudaStream_t)stream>>>(ext_theta_own, ext_pi_own, plane_stride, n, a);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_slab_select(double* theta_own, const double* ext_theta_own, size_t plane_stride, int Lx, int L,
                   const int32_t* accept, u1_stream_t stream) {
    g_launches = 0;
    if (!theta_own || !ext_theta_own || !accept || Lx <= 0 || L <= 0 || plane_stride < (size_t)Lx * L) return U1_ERR_ARG;
    const size_t n = (size_t)Lx * L;
    k_slab_select_planes<<<slab_grid(2 * n), 256, 0, (cudaStream_t)stream>>>(theta_own, ext_theta_own, plane_stride, n, accept);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_slab_select_state(double* state_own, const double* src_own, size_t plane_stride, double* theta_own, int Lx, int L,
                         const int32_t* accept, double* sums_state, const double* sums_new, u1_stream_t stream) {
    g_launches = 0;
    if (!state_own || !src_own || !accept || Lx <= 0 || L <= 0 || plane_stride < (size_t)Lx * L) return U1_ERR_ARG;
    if ((sums_state == nullptr) != (sums_new == nullptr)) return U1_ERR_ARG;
    const size_t n = (size_t)Lx * L;
    k_slab_select_state<<<slab_grid(2 * n), 256, 0, (cudaStream_t)stream>>>(state_own, src_own, plane_stride, n, theta_own, accept,
                                                                            sums_state, sums_new);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_slab_copy_in(const double* theta_own, double* ext_theta_own, size_t plane_stride, int Lx, int L, u1_stream_t stream) {
    g_launches = 0;
    if (!theta_own || !ext_theta_own || Lx <= 0 || L <= 0 || plane_stride < (size_t)Lx * L) return U1_ERR_ARG;
    const size_t n = (size_t)Lx * L;
    k_slab_copy_in<<<slab_grid(2 * n), 256, 0, (cudaStream_t)stream>>>(theta_own, ext_theta_own, plane_stride, n);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

}  // extern "C"
