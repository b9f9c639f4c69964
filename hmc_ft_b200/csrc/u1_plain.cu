// Plain 2D U(1) HMC on sm_100a: streaming (one fused drift+force+kick kernel per leapfrog
// step, HBM-bound, any L / slabs) and SM-resident (whole trajectory in registers, one CTA per
// chain, L in {8,16,32,64}) paths behind the C ABI of include/u1hmc.h.
//
// Reference semantics: 2d_u1_cluster/hmc_u1.py:56-97,219-225 and utils.py:118-196.
#include "u1_common.cuh"
#include "u1_fastmath.cuh"
#include <cstdlib>
#include <atomic>

namespace u1 {

thread_local long long g_launches = 0;

// ---------------------------------------------------------------------------------------
// Elementwise primitives
// ---------------------------------------------------------------------------------------
__global__ void k_plaquette(const double* __restrict__ th, double* __restrict__ out, int B, int L) {
    const long long n = (long long)B * L * L;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n;
         s += (long long)gridDim.x * blockDim.x) {
        const int V = L * L;
        const int b = (int)(s / V), r = (int)(s - (long long)b * V);
        const int x = r / L, y = r - x * L;
        const double* t0 = th + (size_t)b * 2 * V;
        const double* t1 = t0 + V;
        const int yp = wrapi(y + 1, L), xp = wrapi(x + 1, L);
        out[s] = t0[x * L + y] - t1[x * L + y] - t0[x * L + yp] + t1[xp * L + y];
    }
}

__global__ void k_rectangle(const double* __restrict__ th, double* __restrict__ out, int B, int L) {
    const long long n = (long long)B * L * L;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n;
         s += (long long)gridDim.x * blockDim.x) {
        const int V = L * L;
        const int b = (int)(s / V), r = (int)(s - (long long)b * V);
        const int x = r / L, y = r - x * L;
        const double* t0 = th + (size_t)b * 2 * V;
        const double* t1 = t0 + V;
        const int y1 = wrapi(y + 1, L), y2 = wrapi(y + 2, L), x1 = wrapi(x + 1, L), x2 = wrapi(x + 2, L);
        // utils.py:137-139, same term order as the reference
        const double r0 = t0[x * L + y] + t0[x1 * L + y] + t1[x2 * L + y] - t0[x1 * L + y1] -
                          t0[x * L + y1] - t1[x * L + y];
        const double r1 = t0[x * L + y] + t1[x1 * L + y] + t1[x1 * L + y1] - t0[x * L + y2] -
                          t1[x * L + y1] - t1[x * L + y];
        out[(size_t)b * 2 * V + r] = r0;
        out[(size_t)b * 2 * V + V + r] = r1;
    }
}

__global__ void k_regularize(const double* __restrict__ in, double* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = wrap_pi(in[i]);
}

// theta_out = regularize(theta + a*pi): last half drift of hmc_u1.py:78-79.
__global__ void k_final_drift(const double* th, const double* __restrict__ pi,
                              double* out, double a, size_t n) {      // out may alias th (in-place slab update)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = wrap_pi(fma(a, pi[i], th[i]));
}

__global__ void k_momenta(double* __restrict__ pi, uint64_t seed, uint64_t chain0, uint64_t traj,
                          int B, int L) {
    const int V = L * L;
    const long long n = (long long)B * V;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n;
         s += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(s / V), site = (int)(s - (long long)b * V);
        double n0, n1;
        philox_site_normals(chain_key(seed, chain0 + b), traj, (uint32_t)site, kLogTab, n0, n1);
        pi[(size_t)b * 2 * V + site] = n0;
        pi[(size_t)b * 2 * V + V + site] = n1;
    }
}

// Single-pass deterministic reduction for kernels of ONE lattice (the slab path): every block stores its NV partial
// values and draws a ticket; the block that draws the last ticket adds all partials (fixed order: thread-strided, then
// block_sum) into out[0..NV) and resets the ticket, so the kernel needs no separate finishing launch and can be
// replayed from a CUDA graph.  `ticket` must be zero before the first launch.  v[] holds the block totals (valid in
// thread 0, as block_sum leaves them); red: NV*32 doubles of shared memory.
template <int NV>
__device__ __forceinline__ void finish_by_last_block(double (&v)[NV], double* __restrict__ partial, unsigned int* ticket,
                                                     double* __restrict__ out, double* red, int nblk, int my_block) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partial[(size_t)my_block * NV + k] = v[k];
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == (unsigned)(nblk - 1)) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double a[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) a[k] = 0.0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x)
#pragma unroll
        for (int k = 0; k < NV; ++k) a[k] += *reinterpret_cast<const volatile double*>(partial + (size_t)i * NV + k);
    block_sum<NV>(a, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) out[k] = a[k];
        *ticket = 0u;
    }
}

// SUM: also sum pi^2 over the rows [own0, own0 + own_rows) of the buffer (the slab's own rows; the halo rows get
// momenta too but belong to the neighbours' Hamiltonian) into pi2_out[0] - H_old's kinetic term without a pass over pi.
template <bool SUM>
__global__ void __launch_bounds__(256)
k_momenta_rows(double* __restrict__ pi, uint64_t seed, uint64_t chain, uint64_t traj,
               const uint64_t* __restrict__ traj_dev, long long x_first, int Lx, int L, int Lxg,
               int own0, int own_rows, double* __restrict__ partial, unsigned int* ticket, double* __restrict__ pi2_out) {
    __shared__ double red[32];
    if (traj_dev) traj += *traj_dev;                  // device-resident trajectory counter (CUDA-graph replay)
    const long long n = (long long)Lx * L;
    const PhiloxKey key = chain_key(seed, chain);
    double acc[1] = {0.0};
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n;
         s += (long long)gridDim.x * blockDim.x) {
        const int xl = (int)(s / L), y = (int)(s - (long long)xl * L);
        long long xg = (x_first + xl) % Lxg;
        if (xg < 0) xg += Lxg;
        double n0, n1;
        philox_site_normals(key, traj, (uint32_t)(xg * L + y), kLogTab, n0, n1);
        pi[s] = n0;
        pi[n + s] = n1;
        if (SUM && xl >= own0 && xl < own0 + own_rows) acc[0] += n0 * n0 + n1 * n1;
    }
    if (SUM) {
        block_sum<1>(acc, red);
        finish_by_last_block<1>(acc, partial, ticket, pi2_out, red, (int)gridDim.x, (int)blockIdx.x);
    }
}

// hmc_u1.py:65-69 closed form.  Not on the trajectory path (the step kernel fuses it).
__global__ void k_force(const double* __restrict__ th, double* __restrict__ F, int B, int L, double beta) {
    const long long n = (long long)B * L * L;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n;
         s += (long long)gridDim.x * blockDim.x) {
        const int V = L * L;
        const int b = (int)(s / V), r = (int)(s - (long long)b * V);
        const int x = r / L, y = r - x * L;
        const double* t0 = th + (size_t)b * 2 * V;
        const double* t1 = t0 + V;
        const int yp = wrapi(y + 1, L), xp = wrapi(x + 1, L), ym = wrapi(y - 1, L), xm = wrapi(x - 1, L);
        const double p = t0[x * L + y] - t1[x * L + y] - t0[x * L + yp] + t1[xp * L + y];
        const double pl = t0[x * L + ym] - t1[x * L + ym] - t0[x * L + y] + t1[xp * L + ym];   // P(x,y-1)
        const double pu = t0[xm * L + y] - t1[xm * L + y] - t0[xm * L + yp] + t1[x * L + y];   // P(x-1,y)
        const double sp = sin(wrap_pi(p));
        F[(size_t)b * 2 * V + r] = beta * (sp - sin(wrap_pi(pl)));
        F[(size_t)b * 2 * V + V + r] = beta * (-sp + sin(wrap_pi(pu)));
    }
}

// ---------------------------------------------------------------------------------------
// Streaming fused leapfrog step:  theta' = theta + drift*pi ;  pi' = pi - kick*F(theta')
// One CTA = TR x TC sites staged in shared memory with a one-site periodic halo.
// Algorithmic HBM traffic: read theta0,theta1,pi0,pi1 + write the same = 64 B / site-update.
// ---------------------------------------------------------------------------------------
struct StepArgs {
    const double* th;  const double* pi;
    double* tho;       double* pio;
    const double* th_up; const double* pi_up;   // slab halos (row x=-1), [2][L]
    const double* th_dn; const double* pi_dn;   // slab halos (row x=Lx), [2][L]
    int Lx, L, TR, TC, tiles_y;
    double beta, drift, kick;
};

constexpr int kStepThreads = 256;
constexpr int kStepMaxUnits = 2;   // double2 units per thread (tile <= 1024 sites)

template <bool SLAB>
__global__ void __launch_bounds__(kStepThreads) k_leapfrog_step(StepArgs a) {
    extern __shared__ double sm[];
    const int TR = a.TR, TC = a.TC, L = a.L, Lx = a.Lx;
    const int W = TC + 2;                       // padded row width of the theta' tiles
    double* t0 = sm;                            // (TR+2) x (TC+2), origin at (-1,-1)
    double* t1 = t0 + (TR + 2) * W;
    double* sp = t1 + (TR + 2) * W;             // (TR+1) x (TC+1), origin at (-1,-1)
    const int WS = TC + 1;

    const int tile = blockIdx.x;
    const int x0 = (tile / a.tiles_y) * TR, y0 = (tile % a.tiles_y) * TC;
    const size_t V = (size_t)Lx * L;
    const size_t base = (size_t)blockIdx.y * 2 * V;
    const double* g0 = a.th + base;  const double* g1 = g0 + V;
    const double* q0 = a.pi + base;  const double* q1 = q0 + V;
    const int tid = threadIdx.x;

    // Phase 1a: interior, vectorised double2, momenta kept in registers for phase 3.
    double2 pr0[kStepMaxUnits], pr1[kStepMaxUnits];
    const int units = TR * (TC >> 1);
#pragma unroll
    for (int k = 0; k < kStepMaxUnits; ++k) {
        const int u = tid + k * kStepThreads;
        if (u < units) {
            const int r = u / (TC >> 1), c = (u - r * (TC >> 1)) << 1;
            const size_t g = (size_t)(x0 + r) * L + (y0 + c);
            const double2 a0 = *reinterpret_cast<const double2*>(g0 + g);
            const double2 a1 = *reinterpret_cast<const double2*>(g1 + g);
            pr0[k] = *reinterpret_cast<const double2*>(q0 + g);
            pr1[k] = *reinterpret_cast<const double2*>(q1 + g);
            const int o = (r + 1) * W + (c + 1);
            t0[o] = fma(a.drift, pr0[k].x, a0.x);  t0[o + 1] = fma(a.drift, pr0[k].y, a0.y);
            t1[o] = fma(a.drift, pr1[k].x, a1.x);  t1[o + 1] = fma(a.drift, pr1[k].y, a1.y);
        }
    }
    // Phase 1b: halo ring (rows -1 and TR over all columns, columns -1 and TC over interior rows).
    const int ring = 2 * W + 2 * TR;
    for (int h = tid; h < ring; h += kStepThreads) {
        int r, c;
        if (h < 2 * W) { r = (h < W) ? -1 : TR; c = (h < W ? h : h - W) - 1; }
        else { const int k = h - 2 * W; r = k >> 1; c = (k & 1) ? TC : -1; }
        const int gy = wrapi(y0 + c, L);
        int gx = x0 + r;
        double v0, v1;
        if (SLAB && gx < 0) {
            v0 = fma(a.drift, a.pi_up[gy], a.th_up[gy]);
            v1 = fma(a.drift, a.pi_up[L + gy], a.th_up[L + gy]);
        } else if (SLAB && gx >= Lx) {
            v0 = fma(a.drift, a.pi_dn[gy], a.th_dn[gy]);
            v1 = fma(a.drift, a.pi_dn[L + gy], a.th_dn[L + gy]);
        } else {
            gx = wrapi(gx, Lx);
            const size_t g = (size_t)gx * L + gy;
            v0 = fma(a.drift, q0[g], g0[g]);
            v1 = fma(a.drift, q1[g], g1[g]);
        }
        t0[(r + 1) * W + (c + 1)] = v0;
        t1[(r + 1) * W + (c + 1)] = v1;
    }
    __syncthreads();
    // Phase 2: sin P on rows -1..TR-1, cols -1..TC-1.
    for (int i = tid; i < (TR + 1) * WS; i += kStepThreads) {
        const int r = i / WS, c = i - r * WS;       // already offset by +1 w.r.t. (-1,-1)
        const double p = t0[r * W + c] - t1[r * W + c] - t0[r * W + c + 1] + t1[(r + 1) * W + c];
        sp[i] = fast_sin(p);
    }
    __syncthreads();
    // Phase 3: kick + write back.
    const double kb = a.kick * a.beta;
    double* o0 = a.tho + base;  double* o1 = o0 + V;
    double* w0 = a.pio + base;  double* w1 = w0 + V;
#pragma unroll
    for (int k = 0; k < kStepMaxUnits; ++k) {
        const int u = tid + k * kStepThreads;
        if (u < units) {
            const int r = u / (TC >> 1), c = (u - r * (TC >> 1)) << 1;
            const size_t g = (size_t)(x0 + r) * L + (y0 + c);
            const int o = (r + 1) * W + (c + 1);
            const int s = (r + 1) * WS + (c + 1);
            const double s00 = sp[s], s01 = sp[s + 1], sl = sp[s - 1];
            const double su0 = sp[s - WS], su1 = sp[s - WS + 1];
            double2 n0, n1;
            n0.x = fma(-kb, s00 - sl, pr0[k].x);   n0.y = fma(-kb, s01 - s00, pr0[k].y);
            n1.x = fma(-kb, su0 - s00, pr1[k].x);  n1.y = fma(-kb, su1 - s01, pr1[k].y);
            *reinterpret_cast<double2*>(o0 + g) = make_double2(t0[o], t0[o + 1]);
            *reinterpret_cast<double2*>(o1 + g) = make_double2(t1[o], t1[o + 1]);
            *reinterpret_cast<double2*>(w0 + g) = n0;
            *reinterpret_cast<double2*>(w1 + g) = n1;
        }
    }
}

static int largest_divisor_leq(int n, int cap, bool even) {
    for (int d = (cap < n ? cap : n); d >= 1; --d)
        if (n % d == 0 && (!even || (d % 2 == 0))) return d;
    return 1;
}

struct Tiling { int TR, TC, tiles_x, tiles_y; size_t smem; };
static Tiling make_tiling(int Lx, int L) {
    Tiling t;
    t.TC = largest_divisor_leq(L, 128, true);
    t.TR = largest_divisor_leq(Lx, 1024 / t.TC > 0 ? 1024 / t.TC : 1, false);
    t.tiles_x = Lx / t.TR;
    t.tiles_y = L / t.TC;
    t.smem = (size_t)(2 * (t.TR + 2) * (t.TC + 2) + (t.TR + 1) * (t.TC + 1)) * sizeof(double);
    return t;
}

static int launch_step(const StepArgs& in, int B, bool slab, cudaStream_t st) {
    StepArgs a = in;
    if (a.L < 2 || (a.L & 1)) return U1_ERR_SHAPE;
    Tiling t = make_tiling(a.Lx, a.L);
    a.TR = t.TR; a.TC = t.TC; a.tiles_y = t.tiles_y;
    const size_t per = (size_t)2 * a.Lx * a.L;
    for (int b0 = 0; b0 < B; b0 += kMaxGridY) {            // chains ride in gridDim.y (<= 65535 per launch)
        const int nb = (B - b0 < kMaxGridY) ? B - b0 : kMaxGridY;
        StepArgs s = a;
        s.th += (size_t)b0 * per; s.pi += (size_t)b0 * per; s.tho += (size_t)b0 * per; s.pio += (size_t)b0 * per;
        dim3 grid(t.tiles_x * t.tiles_y, nb);
        if (slab) k_leapfrog_step<true><<<grid, kStepThreads, t.smem, st>>>(s);
        else      k_leapfrog_step<false><<<grid, kStepThreads, t.smem, st>>>(s);
        count_launch();
        U1_CHECK_LAUNCH();
    }
    return U1_OK;
}


// ---------------------------------------------------------------------------------------
// Temporally blocked leapfrog for lattices that do not fit one SM (L > 64, slabs): a CTA loads a
// 64 x 64 window of (theta, pi) into registers exactly like k_traj_resident, advances it `ksteps`
// (<= 4) leapfrog steps treating the window as periodic, and writes back only the inner 56 x 56
// sites.  The wrap-around contaminates one ring of sites per step, which never reaches the inner
// block, so the result is bit-identical to `ksteps` launches of k_leapfrog_step while HBM is
// touched once per 4 steps: algorithmic traffic 64 B per site-update, real traffic
// (64^2 + 56^2)/(56^2 * 4) * 32 B = 18.4 B.  Needs L % 4 == 0 (16-byte aligned thread tiles).
// ---------------------------------------------------------------------------------------
constexpr int kMsE = 64, kMsHalo = 4, kMsOut = kMsE - 2 * kMsHalo;   // window rows, halo, inner block rows

struct MultiArgs {
    const double* th; const double* pi;
    double* tho; double* pio;
    int Lx, L, tiles_y, ksteps;
    int tile_x0;          // first tile row of this launch (launches may cover a sub-range of tile rows)
    int split_rows, split_jump;   // tile rows >= split_rows of the launch are shifted by split_jump (two ranges in one launch)
    double beta, first_drift, dt;
};

// Measured and rejected in round 2 (profiles/r02_multistep_prefetch_variants.log): a persistent form (2 CTAs per SM
// walking windows) that asks L2 for the CTA's next window right after a window sits in registers, with
// cp.async.bulk.prefetch.L2 (SASS UBLKPF.L2) or per-line prefetch.global.L2 (CCTL.E.PF2): bit-identical, but 338 / 332 us
// per 4 steps at L=4096 against 307 us for one window per CTA - the hardware CTA scheduler overlaps the load phase of
// a fresh CTA with its neighbour's steps better than the in-kernel loop does.
template <int EW>      // window columns (64: one 512-thread CTA per SM; 32: two 256-thread CTAs per SM overlap their phases)
__global__ void __launch_bounds__((kMsE / 2) * (EW / 4), 512 / ((kMsE / 2) * (EW / 4)))
k_leapfrog_multi(MultiArgs a) {
    constexpr int L = kMsE, TX = 2, TY = 4, NTX = L / TX, NTY = EW / TY, OW = EW - 2 * kMsHalo;
    // edge arrays are [i or j][tx][ty]: the 16 lanes of a half-warp (2 tx x 8 ty at EW=32) hit 16 distinct banks
    extern __shared__ __align__(16) double dsm[];
    double* colA = dsm;
    double* rowA = colA + L * NTY;
    double* colB = rowA + NTX * TY * NTY;
    double* rowB = colB + L * NTY;

    const int tid = threadIdx.x, ty = tid % NTY, tx = tid / NTY;
    const int x0 = tx * TX, y0 = ty * TY;
    const int tyR = (ty + 1) % NTY, tyL = (ty + NTY - 1) % NTY;
    const int txD = (tx + 1) % NTX, txU = (tx + NTX - 1) % NTX;
    const size_t V = (size_t)a.Lx * a.L;
    const double kb = a.dt * a.beta;
    const int tile_xl = blockIdx.x / a.tiles_y, tile_y = blockIdx.x - tile_xl * a.tiles_y;
    const int tile_x = a.tile_x0 + tile_xl + (tile_xl >= a.split_rows ? a.split_jump : 0);
    const size_t base = (size_t)blockIdx.y * 2 * V;
    // global coordinates of this thread's 2 x 4 sites (periodic)
    int gy = tile_y * OW - kMsHalo + y0;
    gy %= a.L; if (gy < 0) gy += a.L;
    int gx[TX];
#pragma unroll
    for (int i = 0; i < TX; ++i) { int g = (tile_x * kMsOut - kMsHalo + x0 + i) % a.Lx; gx[i] = g < 0 ? g + a.Lx : g; }

    double th0[TX][TY], th1[TX][TY], p0[TX][TY], p1[TX][TY];
#pragma unroll
    for (int i = 0; i < TX; ++i)
#pragma unroll
        for (int j = 0; j < TY; j += 2) {
            const size_t g = base + (size_t)gx[i] * a.L + gy + j;
            const double2 v0 = *reinterpret_cast<const double2*>(a.th + g);
            const double2 v1 = *reinterpret_cast<const double2*>(a.th + g + V);
            const double2 q0 = *reinterpret_cast<const double2*>(a.pi + g);
            const double2 q1 = *reinterpret_cast<const double2*>(a.pi + g + V);
            th0[i][j] = v0.x; th0[i][j + 1] = v0.y; th1[i][j] = v1.x; th1[i][j + 1] = v1.y;
            p0[i][j] = q0.x; p0[i][j + 1] = q0.y; p1[i][j] = q1.x; p1[i][j + 1] = q1.y;
        }
    for (int s = 0; s < a.ksteps; ++s) {
        const double dr = (s == 0) ? a.first_drift : a.dt;
#pragma unroll
        for (int i = 0; i < TX; ++i)
#pragma unroll
            for (int j = 0; j < TY; ++j) {
                th0[i][j] = fma(dr, p0[i][j], th0[i][j]);
                th1[i][j] = fma(dr, p1[i][j], th1[i][j]);
            }
        // edges as double2: column pair at [tx][ty], row quadruple as two double2 at [jp][tx][ty] (lanes 16 B apart)
        reinterpret_cast<double2*>(colA)[tx * NTY + ty] = make_double2(th0[0][0], th0[1][0]);
        reinterpret_cast<double2*>(rowA)[(0 * NTX + tx) * NTY + ty] = make_double2(th1[0][0], th1[0][1]);
        reinterpret_cast<double2*>(rowA)[(1 * NTX + tx) * NTY + ty] = make_double2(th1[0][2], th1[0][3]);
        __syncthreads();
        double S[TX][TY];
        {
            double rc[TX], rr[TY];
            {
                const double2 c = reinterpret_cast<const double2*>(colA)[tx * NTY + tyR];
                const double2 q0 = reinterpret_cast<const double2*>(rowA)[(0 * NTX + txD) * NTY + ty];
                const double2 q1 = reinterpret_cast<const double2*>(rowA)[(1 * NTX + txD) * NTY + ty];
                rc[0] = c.x; rc[1] = c.y; rr[0] = q0.x; rr[1] = q0.y; rr[2] = q1.x; rr[3] = q1.y;
            }
#pragma unroll
            for (int i = 0; i < TX; ++i)
#pragma unroll
                for (int j = 0; j < TY; ++j) {
                    const double right = (j + 1 < TY) ? th0[i][j + 1] : rc[i];
                    const double down = (i + 1 < TX) ? th1[i + 1][j] : rr[j];
                    S[i][j] = fast_sin(th0[i][j] - th1[i][j] - right + down);
                }
        }
        reinterpret_cast<double2*>(colB)[tx * NTY + ty] = make_double2(S[0][TY - 1], S[1][TY - 1]);
        reinterpret_cast<double2*>(rowB)[(0 * NTX + tx) * NTY + ty] = make_double2(S[TX - 1][0], S[TX - 1][1]);
        reinterpret_cast<double2*>(rowB)[(1 * NTX + tx) * NTY + ty] = make_double2(S[TX - 1][2], S[TX - 1][3]);
        __syncthreads();
        double lc[TX], ur[TY];
        {
            const double2 c = reinterpret_cast<const double2*>(colB)[tx * NTY + tyL];
            const double2 q0 = reinterpret_cast<const double2*>(rowB)[(0 * NTX + txU) * NTY + ty];
            const double2 q1 = reinterpret_cast<const double2*>(rowB)[(1 * NTX + txU) * NTY + ty];
            lc[0] = c.x; lc[1] = c.y; ur[0] = q0.x; ur[1] = q0.y; ur[2] = q1.x; ur[3] = q1.y;
        }
#pragma unroll
        for (int i = 0; i < TX; ++i)
#pragma unroll
            for (int j = 0; j < TY; ++j) {
                const double left = (j > 0) ? S[i][j - 1] : lc[i];
                const double up = (i > 0) ? S[i - 1][j] : ur[j];
                p0[i][j] = fma(-kb, S[i][j] - left, p0[i][j]);
                p1[i][j] = fma(-kb, up - S[i][j], p1[i][j]);
            }
    }
    // write back the inner block (clipped to the lattice for the last tile row / column)
    const int ly = y0 - kMsHalo;                       // column inside the inner block (multiple of 4)
    if (ly < 0 || ly >= OW || tile_y * OW + ly >= a.L) return;
#pragma unroll
    for (int i = 0; i < TX; ++i) {
        const int lx = x0 + i - kMsHalo;
        if (lx < 0 || lx >= kMsOut || tile_x * kMsOut + lx >= a.Lx) continue;
#pragma unroll
        for (int j = 0; j < TY; j += 2) {
            const size_t g = base + (size_t)gx[i] * a.L + gy + j;
            *reinterpret_cast<double2*>(a.tho + g) = make_double2(th0[i][j], th0[i][j + 1]);
            *reinterpret_cast<double2*>(a.tho + g + V) = make_double2(th1[i][j], th1[i][j + 1]);
            *reinterpret_cast<double2*>(a.pio + g) = make_double2(p0[i][j], p0[i][j + 1]);
            *reinterpret_cast<double2*>(a.pio + g + V) = make_double2(p1[i][j], p1[i][j + 1]);
        }
    }
}

// Measured and rejected in round 2 (profiles/r02_multistep_cluster_dsmem_variant.log, commit c0f5151 has the code): a
// thread-block-cluster form (2/4/8 CTAs side by side share one 64 x 32CY window, column edges pushed into the neighbour
// CTA through distributed shared memory, barrier.cluster arrive/wait with the neighbour-independent work in between).
// Bit-identical and 1.31/1.22/1.18x instead of 1.52x redundant halo work, but 462/486/537 us per 4 steps at L=4096
// against 307 us: each cluster barrier costs ~0.6 us of arrival skew between CTAs on different SMs.

// SM count of the current device (cached per device)
static int sm_count() {
    static std::atomic<int> sms_of[kMaxDevices];
    const int dev = current_device();
    int n = sms_of[dev].load(std::memory_order_acquire);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        sms_of[dev].store(n, std::memory_order_release);
    }
    return n;
}

static bool multi_supported(int Lx, int L) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("U1_MULTISTEP"); enabled = e ? atoi(e) : 1; }
    return enabled && (L % 4 == 0) && Lx >= kMsE && L >= kMsE;
}

static int launch_multi(const double* th, const double* pi, double* tho, double* pio, int B, int Lx, int L,
                        double beta, double first_drift, double dt, int ksteps, cudaStream_t st,
                        int tile_row0 = 0, int tile_rows = -1, int tile_row1 = 0, int tile_rows1 = 0) {
    MultiArgs a{};
    a.th = th; a.pi = pi; a.tho = tho; a.pio = pio;
    a.Lx = Lx; a.L = L; a.ksteps = ksteps; a.beta = beta; a.first_drift = first_drift; a.dt = dt;
    static int ew = 0;
    if (ew == 0) { const char* e = getenv("U1_MULTISTEP_EW"); ew = (e && atoi(e) == 64) ? 64 : 32; }
    const int tiles_all = (Lx + kMsOut - 1) / kMsOut;
    if (tile_rows < 0) tile_rows = tiles_all - tile_row0;
    if (tile_row0 < 0 || tile_rows < 1 || tile_row0 + tile_rows > tiles_all) return U1_ERR_ARG;
    if (tile_rows1 < 0 || (tile_rows1 > 0 && (tile_row1 < tile_row0 + tile_rows || tile_row1 + tile_rows1 > tiles_all))) return U1_ERR_ARG;
    a.tile_x0 = tile_row0;
    a.split_rows = tile_rows;                                  // second range [tile_row1, tile_row1 + tile_rows1) follows the first
    a.split_jump = tile_rows1 > 0 ? tile_row1 - (tile_row0 + tile_rows) : 0;
    const int tiles_x = tile_rows + tile_rows1;
    constexpr int NTX = kMsE / 2;
    if (ew == 64) {
        constexpr int NTY = 16;
        constexpr size_t smem = (size_t)(2 * (kMsE * NTY + NTX * 4 * NTY)) * sizeof(double);
        a.tiles_y = (L + 56 - 1) / 56;
        static std::atomic<bool> configured[kMaxDevices];      // function attributes are per device
        const int dev = current_device();
        if (!configured[dev].load(std::memory_order_acquire)) {
            cudaError_t e = cudaFuncSetAttribute(k_leapfrog_multi<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            configured[dev].store(true, std::memory_order_release);
        }
        for (int b0 = 0; b0 < B; b0 += kMaxGridY) {
            const int nb = (B - b0 < kMaxGridY) ? B - b0 : kMaxGridY;
            MultiArgs s = a;
            const size_t off = (size_t)b0 * 2 * Lx * L;
            s.th += off; s.pi += off; s.tho += off; s.pio += off;
            k_leapfrog_multi<64><<<dim3(tiles_x * a.tiles_y, nb), 512, smem, st>>>(s);
            count_launch();
            U1_CHECK_LAUNCH();
        }
    } else {
        constexpr int NTY = 8;
        constexpr size_t smem = (size_t)(2 * (kMsE * NTY + NTX * 4 * NTY)) * sizeof(double);
        a.tiles_y = (L + 24 - 1) / 24;
        for (int b0 = 0; b0 < B; b0 += kMaxGridY) {
            const int nb = (B - b0 < kMaxGridY) ? B - b0 : kMaxGridY;
            MultiArgs s = a;
            const size_t off = (size_t)b0 * 2 * Lx * L;
            s.th += off; s.pi += off; s.tho += off; s.pio += off;
            k_leapfrog_multi<32><<<dim3(tiles_x * a.tiles_y, nb), 256, smem, st>>>(s);
            count_launch();
            U1_CHECK_LAUNCH();
        }
    }
    return U1_OK;
}

// n leapfrog steps (first drift `first_drift`, then dt) ping-ponging between (thA,piA) and (thB,piB),
// starting from `th_in/pi_in` (which may be a third buffer).  Uses the temporally blocked kernel in
// groups of up to 4 steps when the lattice allows it.  Returns the buffers holding the result.
static int leapfrog_steps(const double* th_in, const double* pi_in, double* thA, double* piA, double* thB,
                          double* piB, int B, int Lx, int L, double beta, double first_drift, double dt, int n,
                          cudaStream_t st, const double** th_res, const double** pi_res) {
    const double* ct = th_in;
    const double* cp = pi_in;
    const bool multi = multi_supported(Lx, L);
    int done = 0, launch = 0;
    while (done < n) {
        const int k = multi ? ((n - done) < kMsHalo ? (n - done) : kMsHalo) : 1;
        double* ot = (ct == thA) ? thB : thA;
        double* op = (ct == thA) ? piB : piA;
        const double fd = (done == 0) ? first_drift : dt;
        if (multi) {
            U1_TRY(launch_multi(ct, cp, ot, op, B, Lx, L, beta, fd, dt, k, st));
        } else {
            StepArgs a{};
            a.th = ct; a.pi = cp; a.tho = ot; a.pio = op;
            a.Lx = Lx; a.L = L; a.beta = beta; a.drift = fd; a.kick = dt;
            U1_TRY(launch_step(a, B, false, st));
        }
        ct = ot; cp = op;
        done += k; ++launch;
    }
    *th_res = ct; *pi_res = cp;
    return U1_OK;
}

// ---------------------------------------------------------------------------------------
// Per-chain sums: [0]=sum cos(wrap P), [1]=sum wrap P, [2]=sum pi^2.  Two deterministic phases.
// ---------------------------------------------------------------------------------------
constexpr int kSumThreads = 256;
constexpr int kSumRowsPerBlock = 16;
constexpr size_t kTicketBytes = 256;      // head of a reduction workspace: the ticket word of finish_by_last_block (zero it once)

__global__ void __launch_bounds__(kSumThreads)
k_partial_sums(const double* __restrict__ th, const double* __restrict__ th_dn,
               const double* __restrict__ pi, double* __restrict__ partial, int Lx, int L, int nblk, int rpb,
               size_t plane, int dn_in_place) {
    // plane = distance between the mu=0 and mu=1 planes (Lx*L for a contiguous field; larger inside a slab's
    // halo-extended buffer, where the row after the last own row IS the lower neighbour's first row).
    __shared__ double red[3 * 32];
    const size_t V = plane;
    const double* t0 = th + (size_t)blockIdx.y * 2 * V;
    const double* t1 = t0 + V;
    const int xa = blockIdx.x * rpb;
    const int xb = min(Lx, xa + rpb);
    double v[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < (xb - xa) * L; i += kSumThreads) {
        const int x = xa + i / L, y = i % L;
        const int yp = wrapi(y + 1, L);
        double t1dn;
        if (x + 1 < Lx || dn_in_place) t1dn = t1[(size_t)(x + 1) * L + y];
        else t1dn = th_dn ? th_dn[L + y] : t1[y];
        // cos(wrap P) = cos P; wrap_fast agrees with utils.py:182-187 to its rounding (u1_fastmath.cuh)
        const double P = t0[(size_t)x * L + y] - t1[(size_t)x * L + y] - t0[(size_t)x * L + yp] + t1dn;
        v[0] += fast_cos(P);
        v[1] += wrap_fast(P);
        if (pi) {
            const double* q = pi + (size_t)blockIdx.y * 2 * V;
            const double a = q[(size_t)x * L + y], b = q[V + (size_t)x * L + y];
            v[2] += a * a + b * b;
        }
    }
    block_sum<3>(v, red);
    if (threadIdx.x == 0) {
        double* o = partial + ((size_t)blockIdx.y * nblk + blockIdx.x) * 3;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    }
}

// Slab path: the last half drift theta' = regularize(theta + a*pi) (hmc_u1.py:78-79) fused into the H_new sums of ONE
// lattice slab in extended layout.  A CTA owns kDsRows rows x 256 columns, one thread per column: all loads of its
// rows are issued up front, every theta' is evaluated once by its owner (the reference's division-based regularize is
// the expensive part: 2 per site + 1 per column for the row below), theta'_0 of the right neighbour comes through
// shared memory (the last thread of the CTA evaluates its neighbour column itself) and theta'_1 of the row below is
// the next row's own value.  theta' of the own sites goes to `tho`; theta AND pi of the row after the last own row
// are read in place.  The sums land in sums[0..2] through finish_by_last_block (no finishing launch).
constexpr int kDsRows = 4;
__global__ void __launch_bounds__(kSumThreads)
k_drift_sums(const double* __restrict__ th, const double* __restrict__ pi, double* __restrict__ tho, size_t plane,
             int Lx, int L, double a, double* __restrict__ partial, unsigned int* ticket, double* __restrict__ sums) {
    __shared__ double red[3 * 32];
    __shared__ double s_a[kDsRows][kSumThreads + 1];
    const int tid = threadIdx.x;
    const int y = blockIdx.x * kSumThreads + tid;
    const int xa = blockIdx.y * kDsRows;
    const bool live = y < L;
    const int yc = live ? y : 0;                               // dead threads shadow column 0
    const bool edge = live && (tid == kSumThreads - 1 || y + 1 == L);   // right neighbour belongs to another CTA / wraps
    const int yp = (yc + 1 == L) ? 0 : yc + 1;
    const double* t0 = th;  const double* t1 = th + plane;
    const double* q0 = pi;  const double* q1 = pi + plane;
    double T0[kDsRows], P0[kDsRows], T1[kDsRows + 1], P1[kDsRows + 1], E0[kDsRows], EP[kDsRows];
#pragma unroll
    for (int r = 0; r <= kDsRows; ++r) {
        const int x = min(xa + r, Lx);                         // row Lx = the lower neighbour's first row (in place)
        const size_t o = (size_t)x * L + yc;
        T1[r] = t1[o]; P1[r] = q1[o];
        if (r < kDsRows) {
            T0[r] = t0[o]; P0[r] = q0[o];
            E0[r] = 0.0; EP[r] = 0.0;
            if (edge) { E0[r] = t0[(size_t)x * L + yp]; EP[r] = q0[(size_t)x * L + yp]; }
        }
    }
    double A[kDsRows], Bv[kDsRows + 1];
#pragma unroll
    for (int r = 0; r <= kDsRows; ++r) {
        Bv[r] = wrap_pi(fma(a, P1[r], T1[r]));
        if (r < kDsRows) {
            A[r] = wrap_pi(fma(a, P0[r], T0[r]));
            if (live) s_a[r][tid] = A[r];                      // (a dead thread's slot is written by its left neighbour: `edge`)
        }
    }
    if (edge) {
#pragma unroll
        for (int r = 0; r < kDsRows; ++r) s_a[r][tid + 1] = wrap_pi(fma(a, EP[r], E0[r]));
    }
    __syncthreads();
    double v[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < kDsRows; ++r) {
        const int x = xa + r;
        if (live && x < Lx) {
            const size_t o = (size_t)x * L + y;
            tho[o] = A[r];
            tho[plane + o] = Bv[r];
            const double P = A[r] - Bv[r] - s_a[r][tid + 1] + Bv[r + 1];
            v[0] += fast_cos(P);
            v[1] += wrap_fast(P);
            v[2] += P0[r] * P0[r] + P1[r] * P1[r];
        }
    }
    block_sum<3>(v, red);
    finish_by_last_block<3>(v, partial, ticket, sums, red, (int)(gridDim.x * gridDim.y), (int)(blockIdx.y * gridDim.x + blockIdx.x));
}

// one CTA per chain: fixed-order (deterministic) reduction of the nblk partial triples
__global__ void __launch_bounds__(128) k_finish_sums(const double* __restrict__ partial, double* __restrict__ sums, int nblk) {
    __shared__ double red[3 * 32];
    const int b = blockIdx.x;
    double v[3] = {0.0, 0.0, 0.0};
    for (int k = threadIdx.x; k < nblk; k += 128) {
        const double* p = partial + ((size_t)b * nblk + k) * 3;
        v[0] += p[0]; v[1] += p[1]; v[2] += p[2];
    }
    block_sum<3>(v, red);
    if (threadIdx.x == 0) {
        sums[(size_t)b * 3 + 0] = v[0]; sums[(size_t)b * 3 + 1] = v[1]; sums[(size_t)b * 3 + 2] = v[2];
    }
}

// Rows per CTA: 16 for big batches; fewer when B*Lx/16 CTAs would not fill the GPU (one lattice slab:
// Lx=512 would be 32 CTAs), down to one row per CTA.
static inline int sum_rows_per_block(int B, int Lx) {
    int rpb = kSumRowsPerBlock;
    while (rpb > 1 && (long long)B * ((Lx + rpb - 1) / rpb) < 148 * 8) rpb >>= 1;
    return rpb;
}
static inline int sum_blocks(int B, int Lx) { const int r = sum_rows_per_block(B, Lx); return (Lx + r - 1) / r; }

// sums: device [B][3]; partial: device [B][nblk][3]
static int launch_sums(const double* th, const double* th_dn, const double* pi, double* sums,
                       double* partial, int B, int Lx, int L, cudaStream_t st, size_t plane = 0, int dn_in_place = 0) {
    const int rpb = sum_rows_per_block(B, Lx), nblk = (Lx + rpb - 1) / rpb;
    const size_t pl = plane ? plane : (size_t)Lx * L;
    for (int b0 = 0; b0 < B; b0 += kMaxGridY) {            // chains ride in gridDim.y (<= 65535 per launch)
        const int nb = (B - b0 < kMaxGridY) ? B - b0 : kMaxGridY;
        k_partial_sums<<<dim3(nblk, nb), kSumThreads, 0, st>>>(th + (size_t)b0 * 2 * pl, th_dn, pi ? pi + (size_t)b0 * 2 * pl : nullptr,
                                                               partial + (size_t)b0 * nblk * 3, Lx, L, nblk, rpb, pl, dn_in_place);
        count_launch();
        U1_CHECK_LAUNCH();
    }
    k_finish_sums<<<B, 128, 0, st>>>(partial, sums, nblk);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

__global__ void k_scale_action(const double* __restrict__ sums, double* __restrict__ S, int B, double beta) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) S[b] = (-beta) * sums[(size_t)b * 3];
}

__global__ void k_obs_from_sums(const double* __restrict__ sums, double* __restrict__ plaq,
                                double* __restrict__ topo, int B, double inv_vol) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (plaq) plaq[b] = sums[(size_t)b * 3] * inv_vol;
    if (topo) topo[b] = floor(0.1 + sums[(size_t)b * 3 + 1] / kTwoPi);
}

// Metropolis decision for the streaming path (hmc_u1.py:84-97).
__global__ void k_metropolis(const double* __restrict__ s_old, const double* __restrict__ s_new,
                             const double* __restrict__ u_inject, uint64_t seed, uint64_t chain0,
                             uint64_t traj, const uint64_t* __restrict__ traj_dev, int B, double beta, double inv_vol,
                             double* out_dH, int32_t* out_acc, double* out_H, double* out_plaq,
                             double* out_topo, int32_t* acc_flag) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (traj_dev) traj += *traj_dev;
    const double* o = s_old + (size_t)b * 3;
    const double* n = s_new + (size_t)b * 3;
    const double H_old = (-beta) * o[0] + 0.5 * o[2];
    const double H_new = (-beta) * n[0] + 0.5 * n[2];
    const double dH = H_new - H_old;
    const double u = u_inject ? u_inject[b] : philox_accept_uniform(chain_key(seed, chain0 + b), traj);
    const bool acc = u < exp(-dH);
    acc_flag[b] = acc ? 1 : 0;
    const double* f = acc ? n : o;
    if (out_dH) out_dH[b] = dH;
    if (out_acc) out_acc[b] = acc ? 1 : 0;
    if (out_H) out_H[b] = acc ? H_new : H_old;
    if (out_plaq) out_plaq[b] = f[0] * inv_vol;
    if (out_topo) out_topo[b] = floor(0.1 + f[1] / kTwoPi);
}

__global__ void k_select(double* __restrict__ theta, const double* __restrict__ theta_new,
                         const int32_t* __restrict__ acc, size_t per_chain, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        if (acc[i / per_chain]) theta[i] = theta_new[i];
}

// ---------------------------------------------------------------------------------------
// SM-resident whole-trajectory kernel.  One CTA per chain; each thread owns a TX x TY tile of
// sites (theta0, theta1, pi0, pi1 in registers for the whole trajectory).  Per leapfrog step
// the only exchanged data are tile edges (theta' first column/row, sinP last column/row)
// through shared memory: two __syncthreads per step, no global traffic.
// ---------------------------------------------------------------------------------------
struct TrajArgs {
    double* theta;
    const double* pi_inject; const double* u_inject;
    uint64_t seed, chain0, traj0;
    int n_traj, B, n_steps;
    double beta, dt;
    double* out_dH; int32_t* out_acc; double* out_H; double* out_plaq; double* out_topo;
};

// PF = persistent CTAs with a TMA bulk prefetch: the grid is one CTA per SM, every CTA walks chains
// b = blockIdx.x, blockIdx.x + gridDim.x, ... and, as soon as its threads hold the current chain in
// registers, one thread starts `cp.async.bulk` (UBLKCP) of the NEXT chain's 2*V doubles into a shared-memory
// stage, completion on an mbarrier; the next trajectory then starts from shared memory instead of
// waiting on HBM with an otherwise idle SM (one CTA fills the register file, so nothing else hides it).
template <int L, int TX, int TY, bool PF>
__global__ void __launch_bounds__((L / TX) * (L / TY))
k_traj_resident(TrajArgs a) {
    constexpr int NTX = L / TX, NTY = L / TY, NT = NTX * NTY, V = L * L;
    static_assert(NT % 32 == 0 && TY % 2 == 0, "tile config");
    extern __shared__ __align__(16) double dsm[];
    double* colA = dsm;                       // theta0' at (x, first column of tile ty): [x][ty]
    double* rowA = colA + L * NTY;            // theta1' at (first row of tile tx, y): [tx][j][ty]
    double* colB = rowA + NTX * TY * NTY;     // sinP at (x, last column of tile ty)
    double* rowB = colB + L * NTY;            // sinP at (last row of tile tx, y)
    double* red = rowB + NTX * TY * NTY;      // 3*32 reduction scratch
    LogTabEntry* ltab = reinterpret_cast<LogTabEntry*>(red + 96);   // Box-Muller log table (2 KB)
    double* stage = reinterpret_cast<double*>(ltab + kLogTabN);       // PF: next chain's theta [2][L][L]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(stage + 2 * V);      // PF: completion barrier of the bulk copy

    const int tid = threadIdx.x, ty = tid % NTY, tx = tid / NTY;
    const int x0 = tx * TX, y0 = ty * TY;
    const int tyR = (ty + 1) % NTY, tyL = (ty + NTY - 1) % NTY;
    const int txD = (tx + 1) % NTX, txU = (tx + NTX - 1) % NTX;
    const double beta = a.beta, dt = a.dt, kb = a.dt * a.beta;

    double th0[TX][TY], th1[TX][TY], p0[TX][TY], p1[TX][TY];
    for (int i = tid; i < kLogTabN; i += NT) ltab[i] = kLogTab[i];
    const unsigned mbar_s = (unsigned)__cvta_generic_to_shared(mbar);
    auto prefetch_chain = [&](int bn) {              // thread 0 only
        constexpr unsigned kBytes = 2u * V * (unsigned)sizeof(double);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(kBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(stage)), "l"(a.theta + (size_t)bn * 2 * V), "r"(kBytes), "r"(mbar_s)
                     : "memory");
    };
    if (PF && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if ((int)blockIdx.x < a.B) prefetch_chain(blockIdx.x);
    }
    unsigned parity = 0;
    __syncthreads();

    // Edge exchange through shared memory.  With 2 x 4 tiles the column edge (2 values) is one double2 at
    // [tx][ty] and the row edge (4 values) two double2 at [jp][tx][ty]: 3 STS.128 + 3 LDS.128 per exchange
    // instead of 6 + 6 64-bit accesses, lanes 16 bytes apart (conflict-free).  Other tile shapes keep the
    // scalar layout.
    constexpr bool VEC = (TX == 2 && TY == 4);
    auto put_edges = [&](double* col, double* row, const double (&cv)[TX], const double (&rv)[TY]) {
        if constexpr (VEC) {
            reinterpret_cast<double2*>(col)[tx * NTY + ty] = make_double2(cv[0], cv[1]);
            reinterpret_cast<double2*>(row)[(0 * NTX + tx) * NTY + ty] = make_double2(rv[0], rv[1]);
            reinterpret_cast<double2*>(row)[(1 * NTX + tx) * NTY + ty] = make_double2(rv[2], rv[3]);
        } else {
#pragma unroll
            for (int i = 0; i < TX; ++i) col[(x0 + i) * NTY + ty] = cv[i];
#pragma unroll
            for (int j = 0; j < TY; ++j) row[(tx * TY + j) * NTY + ty] = rv[j];
        }
    };
    auto get_edges = [&](const double* col, const double* row, int tyN, int txN, double (&cv)[TX], double (&rv)[TY]) {
        if constexpr (VEC) {
            const double2 c = reinterpret_cast<const double2*>(col)[tx * NTY + tyN];
            const double2 r0 = reinterpret_cast<const double2*>(row)[(0 * NTX + txN) * NTY + ty];
            const double2 r1 = reinterpret_cast<const double2*>(row)[(1 * NTX + txN) * NTY + ty];
            cv[0] = c.x; cv[1] = c.y; rv[0] = r0.x; rv[1] = r0.y; rv[2] = r1.x; rv[3] = r1.y;
        } else {
#pragma unroll
            for (int i = 0; i < TX; ++i) cv[i] = col[(x0 + i) * NTY + tyN];
#pragma unroll
            for (int j = 0; j < TY; ++j) rv[j] = row[(txN * TY + j) * NTY + ty];
        }
    };
    auto publish_theta = [&]() {
        double cv[TX], rv[TY];
#pragma unroll
        for (int i = 0; i < TX; ++i) cv[i] = th0[i][0];
#pragma unroll
        for (int j = 0; j < TY; ++j) rv[j] = th1[0][j];
        put_edges(colA, rowA, cv, rv);
    };
    auto plaquettes = [&](double (&P)[TX][TY]) {   // call after publish_theta + __syncthreads
        double rc[TX], rr[TY];
        get_edges(colA, rowA, tyR, txD, rc, rr);
#pragma unroll
        for (int i = 0; i < TX; ++i)
#pragma unroll
            for (int j = 0; j < TY; ++j) {
                const double right = (j + 1 < TY) ? th0[i][j + 1] : rc[i];
                const double down = (i + 1 < TX) ? th1[i + 1][j] : rr[j];
                P[i][j] = th0[i][j] - th1[i][j] - right + down;
            }
    };
    // v = {sum cos P, -(sum of plaquette winding numbers) = topological charge, sum pi^2}.  cos(wrap P) =
    // cos P and sum wrap(P) = -2 pi sum n(P) (utils.py:173-196), so neither the wrap nor cos()'s
    // general-range reduction is needed: fast_cos + an integer winding count.
    auto hamiltonian_sums = [&](double (&v)[3]) {
        double P[TX][TY];
        publish_theta();
        __syncthreads();
        plaquettes(P);
        v[0] = v[2] = 0.0;
        int nw = 0;
#pragma unroll
        for (int i = 0; i < TX; ++i)
#pragma unroll
            for (int j = 0; j < TY; ++j) {
                v[0] += fast_cos(P[i][j]);
                nw += winding(P[i][j]);
                v[2] += p0[i][j] * p0[i][j] + p1[i][j] * p1[i][j];
            }
        v[1] = (double)(-nw);
        block_sum<3>(v, red);
    };

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    double* g0 = a.theta + (size_t)b * 2 * V;
    double* g1 = g0 + V;
    const PhiloxKey key = chain_key(a.seed, a.chain0 + b);
    for (int t = 0; t < a.n_traj; ++t) {
        const uint64_t traj = a.traj0 + (uint64_t)t;
        // ---- load state, draw momenta ----
        const bool staged = PF && t == 0;
        if (staged) {                                 // wait for the bulk copy of this chain
            asm volatile("{\n.reg .pred P1;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}"
                         ::"r"(mbar_s), "r"(parity) : "memory");
            parity ^= 1u;
        }
        const double* s0 = staged ? stage : g0;
        const double* s1 = staged ? stage + V : g1;
#pragma unroll
        for (int i = 0; i < TX; ++i)
#pragma unroll
            for (int j = 0; j < TY; j += 2) {
                const int g = (x0 + i) * L + y0 + j;
                const double2 v0 = *reinterpret_cast<const double2*>(s0 + g);
                const double2 v1 = *reinterpret_cast<const double2*>(s1 + g);
                th0[i][j] = v0.x; th0[i][j + 1] = v0.y;
                th1[i][j] = v1.x; th1[i][j + 1] = v1.y;
            }
        if (staged) {                                 // every thread has its sites: the stage is free for the next chain
            __syncthreads();
            if (tid == 0 && b + (int)gridDim.x < a.B) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                prefetch_chain(b + gridDim.x);
            }
        }
        if (a.pi_inject) {
            const double* q0 = a.pi_inject + ((size_t)t * a.B + b) * 2 * V;
            const double* q1 = q0 + V;
#pragma unroll
            for (int i = 0; i < TX; ++i)
#pragma unroll
                for (int j = 0; j < TY; ++j) {
                    p0[i][j] = q0[(x0 + i) * L + y0 + j];
                    p1[i][j] = q1[(x0 + i) * L + y0 + j];
                }
        } else {
#pragma unroll
            for (int i = 0; i < TX; ++i)
#pragma unroll
                for (int j = 0; j < TY; ++j)
                    philox_site_normals(key, traj, (uint32_t)((x0 + i) * L + y0 + j), ltab, p0[i][j], p1[i][j]);
        }
        // ---- H_old (hmc_u1.py:84-85) ----
        double so[3];
        hamiltonian_sums(so);
        const double H_old = (-beta) * so[0] + 0.5 * so[2];

        // ---- leapfrog (hmc_u1.py:71-80) ----
        for (int s = 0; s < a.n_steps; ++s) {
            const double dr = (s == 0) ? 0.5 * dt : dt;
#pragma unroll
            for (int i = 0; i < TX; ++i)
#pragma unroll
                for (int j = 0; j < TY; ++j) {
                    th0[i][j] = fma(dr, p0[i][j], th0[i][j]);
                    th1[i][j] = fma(dr, p1[i][j], th1[i][j]);
                }
            publish_theta();
            __syncthreads();
            double S[TX][TY];
            plaquettes(S);
#pragma unroll
            for (int i = 0; i < TX; ++i)
#pragma unroll
                for (int j = 0; j < TY; ++j) S[i][j] = fast_sin(S[i][j]);
            {
                double cv[TX], rv[TY];
#pragma unroll
                for (int i = 0; i < TX; ++i) cv[i] = S[i][TY - 1];
#pragma unroll
                for (int j = 0; j < TY; ++j) rv[j] = S[TX - 1][j];
                put_edges(colB, rowB, cv, rv);
            }
            __syncthreads();
            double lc[TX], ur[TY];
            get_edges(colB, rowB, tyL, txU, lc, ur);
#pragma unroll
            for (int i = 0; i < TX; ++i)
#pragma unroll
                for (int j = 0; j < TY; ++j) {
                    const double left = (j > 0) ? S[i][j - 1] : lc[i];
                    const double up = (i > 0) ? S[i - 1][j] : ur[j];
                    p0[i][j] = fma(-kb, S[i][j] - left, p0[i][j]);
                    p1[i][j] = fma(-kb, up - S[i][j], p1[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < TX; ++i)
#pragma unroll
            for (int j = 0; j < TY; ++j) {
                th0[i][j] = wrap_fast(fma(0.5 * dt, p0[i][j], th0[i][j]));
                th1[i][j] = wrap_fast(fma(0.5 * dt, p1[i][j], th1[i][j]));
            }
        // ---- H_new, accept/reject (hmc_u1.py:88-97) ----
        double sn[3];
        hamiltonian_sums(sn);
        const double H_new = (-beta) * sn[0] + 0.5 * sn[2];
        const double dH = H_new - H_old;
        const double u = a.u_inject ? a.u_inject[(size_t)t * a.B + b] : philox_accept_uniform(key, traj);
        const bool acc = u < exp(-dH);
        if (acc) {
#pragma unroll
            for (int i = 0; i < TX; ++i)
#pragma unroll
                for (int j = 0; j < TY; j += 2) {
                    const int g = (x0 + i) * L + y0 + j;
                    *reinterpret_cast<double2*>(g0 + g) = make_double2(th0[i][j], th0[i][j + 1]);
                    *reinterpret_cast<double2*>(g1 + g) = make_double2(th1[i][j], th1[i][j + 1]);
                }
        }
        if (tid == 0) {
            const size_t o = (size_t)t * a.B + b;
            const double* f = acc ? sn : so;
            if (a.out_dH) a.out_dH[o] = dH;
            if (a.out_acc) a.out_acc[o] = acc ? 1 : 0;
            if (a.out_H) a.out_H[o] = acc ? H_new : H_old;
            if (a.out_plaq) a.out_plaq[o] = f[0] * (1.0 / V);
            if (a.out_topo) a.out_topo[o] = f[1];
        }
    }
    }
}

// ---- training-loss reduction (field_trans_opt.py:452-458): sum |a-b|^p for p = 2,4,6,8 over n elements ----
__global__ void __launch_bounds__(256) k_diff_pow_sums(const double* __restrict__ a, const double* __restrict__ b, size_t n,
                                                       double* __restrict__ partial) {
    __shared__ double red[4 * 32];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double d = a[i] - b[i], d2 = d * d, d4 = d2 * d2;
        v[0] += d2; v[1] += d4; v[2] += d4 * d2; v[3] += d4 * d4;
    }
    block_sum<4>(v, red);
    if (threadIdx.x == 0) {
        double* o = partial + (size_t)blockIdx.x * 4;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3];
    }
}
__global__ void __launch_bounds__(128) k_finish_pow_sums(const double* __restrict__ partial, int nblk, double* __restrict__ out) {
    __shared__ double red[4 * 32];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = threadIdx.x; k < nblk; k += 128) {
        const double* p = partial + (size_t)k * 4;
        v[0] += p[0]; v[1] += p[1]; v[2] += p[2]; v[3] += p[3];
    }
    block_sum<4>(v, red);
    if (threadIdx.x == 0) { out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; out[3] = v[3]; }
}

__global__ void __launch_bounds__(256) k_fp64_probe(double* __restrict__ out, int iters) {
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
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

// ---- observable post-processing (utils.py:224-298), one CTA per (lag, chain) ----
__global__ void __launch_bounds__(256) k_series_stats(const double* __restrict__ topo, int T, int B, double* __restrict__ stats) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    double s[1] = {0.0};
    for (int t = threadIdx.x; t < T; t += blockDim.x) s[0] += rint(topo[(size_t)t * B + b]);
    block_sum<1>(s, red);
    const double mean = s[0] / T;
    double v[1] = {0.0};
    for (int t = threadIdx.x; t < T; t += blockDim.x) { const double d = rint(topo[(size_t)t * B + b]) - mean; v[0] += d * d; }
    block_sum<1>(v, red);
    if (threadIdx.x == 0) { stats[2 * b] = mean; stats[2 * b + 1] = v[0] / T; }
}

__global__ void __launch_bounds__(256) k_autocorr(const double* __restrict__ topo, int T, int B, double two_v_chi,
                                                   const double* __restrict__ stats, double* __restrict__ gchi,
                                                   double* __restrict__ gdef) {
    __shared__ double red[2 * 32];
    const int d = blockIdx.x, b = blockIdx.y;
    const double mean = stats[2 * b], var = stats[2 * b + 1];
    double v[2] = {0.0, 0.0};
    for (int t = threadIdx.x; t + d < T; t += blockDim.x) {
        const double q0 = rint(topo[(size_t)t * B + b]), q1 = rint(topo[(size_t)(t + d) * B + b]);
        v[0] += (q0 - q1) * (q0 - q1);
        v[1] += (q0 - mean) * (q1 - mean);
    }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) {
        const size_t o = (size_t)d * B + b;
        const int n = T - d;
        if (gchi) gchi[o] = (d == 0) ? 1.0 : 1.0 - (v[0] / n) / two_v_chi;
        if (gdef) gdef[o] = (d == 0 || var == 0.0) ? 1.0 : (v[1] / n) / var;
    }
}

static bool resident_supported(int L) { return L == 8 || L == 16 || L == 32 || L == 64; }

template <int L, int TX, int TY, bool PF>
static int launch_resident_pf(const TrajArgs& a, cudaStream_t st) {
    constexpr int NTX = L / TX, NTY = L / TY;
    constexpr size_t smem = (size_t)(2 * (L * NTY + NTX * TY * NTY) + 3 * 32) * sizeof(double) + kLogTabN * sizeof(LogTabEntry) +
                            (PF ? (size_t)2 * L * L * sizeof(double) + 16 : 0);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_traj_resident<L, TX, TY, PF>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    int grid = a.B;
    if (PF) {                                         // persistent: as many CTAs as can be resident at once
        static std::atomic<int> slots_of[kMaxDevices];        // per device, race-free (same value from any thread)
        const int dev = current_device();
        int slots = slots_of[dev].load(std::memory_order_acquire);
        if (slots == 0) {
            int sms = 0, occ = 0;
            if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_traj_resident<L, TX, TY, PF>, NTX * NTY, smem) != cudaSuccess || occ < 1) occ = 1;
            slots = sms * occ;
            slots_of[dev].store(slots, std::memory_order_release);
        }
        if (grid > slots) grid = slots;
    }
    k_traj_resident<L, TX, TY, PF><<<grid, NTX * NTY, smem, st>>>(a);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}
template <int L, int TX, int TY>
static int launch_resident_t(const TrajArgs& a, cudaStream_t st) {
    static int pf = -1;                               // U1_RESIDENT_PF=0: one CTA per chain, loads from global
    if (pf < 0) { const char* e = getenv("U1_RESIDENT_PF"); pf = e ? atoi(e) : 1; }
    return pf ? launch_resident_pf<L, TX, TY, true>(a, st) : launch_resident_pf<L, TX, TY, false>(a, st);
}

static int launch_resident(const TrajArgs& a, int L, cudaStream_t st) {
    switch (L) {
        case 8:  return launch_resident_t<8, 1, 2>(a, st);
        case 16: return launch_resident_t<16, 2, 4>(a, st);
        case 32: return launch_resident_t<32, 2, 4>(a, st);
        case 64: return (getenv("U1_RESIDENT_TX1") ? launch_resident_t<64, 1, 4>(a, st) : launch_resident_t<64, 2, 4>(a, st));
        default: return U1_ERR_UNSUPPORTED;
    }
}

static inline int ew_grid(size_t n, int threads = 256) {
    size_t g = (n + threads - 1) / threads;
    const size_t cap = 148 * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Workspace layout of the streaming path.
struct PlainWs {
    double *thA, *piA, *thB, *piB, *partial, *s_old, *s_new;
    int32_t* acc;
    size_t bytes;
};
static PlainWs carve_plain(void* ws, int B, int L) {
    PlainWs w;
    const size_t f = align_up((size_t)B * 2 * L * L * sizeof(double), 256);
    const size_t part = align_up((size_t)B * sum_blocks(B, L) * 3 * sizeof(double), 256);
    const size_t s3 = align_up((size_t)B * 3 * sizeof(double), 256);
    char* p = (char*)ws;
    w.thA = (double*)p; p += f;  w.piA = (double*)p; p += f;
    w.thB = (double*)p; p += f;  w.piB = (double*)p; p += f;
    w.partial = (double*)p; p += part;
    w.s_old = (double*)p; p += s3;  w.s_new = (double*)p; p += s3;
    w.acc = (int32_t*)p; p += align_up((size_t)B * sizeof(int32_t), 256);
    w.bytes = (size_t)(p - (char*)ws);
    return w;
}

}  // namespace u1

using namespace u1;

extern "C" {

int u1_abi_version(void) { return U1_ABI_VERSION; }

const char* u1_status_string(int s) {
    switch (s) {
        case U1_OK: return "ok";
        case U1_ERR_ARG: return "invalid argument (null pointer or non-positive size)";
        case U1_ERR_SHAPE: return "unsupported lattice shape (L must be even and >= 2)";
        case U1_ERR_WORKSPACE: return "workspace too small or misaligned";
        case U1_ERR_HANDLE: return "bad FT handle or weight blob";
        case U1_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return s > 0 ? cudaGetErrorString((cudaError_t)s) : "unknown status";
    }
}

int u1_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return (int)e;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return U1_OK;
}

long long u1_last_launch_count(void) { return g_launches; }

int u1_fp64_probe(double* out, int blocks, int iters, u1_stream_t stream) {
    g_launches = 0;
    if (!out || blocks <= 0 || iters <= 0) return U1_ERR_ARG;
    k_fp64_probe<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_autocorr(const double* topo, int T, int B, int max_lag, double two_v_chi,
                double* gamma_chi, double* gamma_def, void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!topo || T < 1 || B < 1 || max_lag < 0 || max_lag >= T || B > 65535) return U1_ERR_ARG;
    if (gamma_chi && !(two_v_chi > 0)) return U1_ERR_ARG;
    if (!ws || ws_bytes < (size_t)2 * B * sizeof(double) || (reinterpret_cast<uintptr_t>(ws) & 7u)) return U1_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    k_series_stats<<<B, 256, 0, st>>>(topo, T, B, (double*)ws);
    count_launch();
    U1_CHECK_LAUNCH();
    k_autocorr<<<dim3(max_lag + 1, B), 256, 0, st>>>(topo, T, B, two_v_chi, (const double*)ws, gamma_chi, gamma_def);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

size_t u1_diff_power_sums_workspace_bytes(void) { return (size_t)148 * 16 * 4 * sizeof(double); }

int u1_diff_power_sums(const double* a, const double* b, size_t n, double* out4, void* ws, size_t ws_bytes,
                       u1_stream_t stream) {
    g_launches = 0;
    if (!a || !b || !out4 || n == 0) return U1_ERR_ARG;
    if (!ws || ws_bytes < u1_diff_power_sums_workspace_bytes() || (reinterpret_cast<uintptr_t>(ws) & 255u)) return U1_ERR_WORKSPACE;
    const int nblk = ew_grid(n);
    k_diff_pow_sums<<<nblk, 256, 0, (cudaStream_t)stream>>>(a, b, n, (double*)ws);
    count_launch();
    U1_CHECK_LAUNCH();
    k_finish_pow_sums<<<1, 128, 0, (cudaStream_t)stream>>>((const double*)ws, nblk, out4);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_plaquette(const double* theta, double* plaq, int B, int L, u1_stream_t stream) {
    g_launches = 0;
    if (!theta || !plaq || B <= 0 || L <= 0) return U1_ERR_ARG;
    k_plaquette<<<ew_grid((size_t)B * L * L), 256, 0, (cudaStream_t)stream>>>(theta, plaq, B, L);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_rectangle(const double* theta, double* rect, int B, int L, u1_stream_t stream) {
    g_launches = 0;
    if (!theta || !rect || B <= 0 || L <= 0) return U1_ERR_ARG;
    k_rectangle<<<ew_grid((size_t)B * L * L), 256, 0, (cudaStream_t)stream>>>(theta, rect, B, L);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_regularize(const double* in, double* out, size_t n, u1_stream_t stream) {
    g_launches = 0;
    if (!in || !out) return U1_ERR_ARG;
    if (n == 0) return U1_OK;
    k_regularize<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(in, out, n);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

size_t u1_workspace_bytes(int B, int L) {
    if (B <= 0 || L <= 0) return 0;
    return carve_plain(nullptr, B, L).bytes;
}

static int check_ws(void* ws, size_t have, size_t need) {
    if (!ws || have < need || (reinterpret_cast<uintptr_t>(ws) & 255u)) return U1_ERR_WORKSPACE;
    return U1_OK;
}

int u1_observables(const double* theta, double* plaq_mean, double* topo, int B, int L,
                   void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!theta || B <= 0 || L <= 0) return U1_ERR_ARG;
    U1_TRY(check_ws(ws, ws_bytes, u1_workspace_bytes(B, L)));
    PlainWs w = carve_plain(ws, B, L);
    cudaStream_t st = (cudaStream_t)stream;
    U1_TRY(launch_sums(theta, nullptr, nullptr, w.s_new, w.partial, B, L, L, st));
    k_obs_from_sums<<<(B + 127) / 128, 128, 0, st>>>(w.s_new, plaq_mean, topo, B, 1.0 / ((double)L * L));
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_action(const double* theta, double* S, int B, int L, double beta,
              void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!theta || !S || B <= 0 || L <= 0) return U1_ERR_ARG;
    U1_TRY(check_ws(ws, ws_bytes, u1_workspace_bytes(B, L)));
    PlainWs w = carve_plain(ws, B, L);
    cudaStream_t st = (cudaStream_t)stream;
    U1_TRY(launch_sums(theta, nullptr, nullptr, w.s_new, w.partial, B, L, L, st));
    k_scale_action<<<(B + 127) / 128, 128, 0, st>>>(w.s_new, S, B, beta);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_force(const double* theta, double* force, int B, int L, double beta, u1_stream_t stream) {
    g_launches = 0;
    if (!theta || !force || B <= 0 || L <= 0) return U1_ERR_ARG;
    k_force<<<ew_grid((size_t)B * L * L), 256, 0, (cudaStream_t)stream>>>(theta, force, B, L, beta);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_momenta(double* pi, uint64_t seed, uint64_t chain0, uint64_t traj, int B, int L,
               u1_stream_t stream) {
    g_launches = 0;
    if (!pi || B <= 0 || L <= 0) return U1_ERR_ARG;
    k_momenta<<<ew_grid((size_t)B * L * L), 256, 0, (cudaStream_t)stream>>>(pi, seed, chain0, traj, B, L);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_leapfrog_step(const double* theta_in, const double* pi_in, double* theta_out, double* pi_out,
                     int B, int L, double beta, double drift, double kick, u1_stream_t stream) {
    g_launches = 0;
    if (!theta_in || !pi_in || !theta_out || !pi_out || B <= 0 || L <= 0) return U1_ERR_ARG;
    if (theta_in == theta_out || pi_in == pi_out) return U1_ERR_ARG;
    if (!aligned16(theta_in) || !aligned16(pi_in) || !aligned16(theta_out) || !aligned16(pi_out)) return U1_ERR_ARG;
    StepArgs a{};
    a.th = theta_in; a.pi = pi_in; a.tho = theta_out; a.pio = pi_out;
    a.Lx = L; a.L = L; a.beta = beta; a.drift = drift; a.kick = kick;
    return launch_step(a, B, false, (cudaStream_t)stream);
}

int u1_leapfrog_step_rect(const double* theta_in, const double* pi_in, double* theta_out, double* pi_out,
                          int B, int Lx, int L, double beta, double drift, double kick,
                          u1_stream_t stream) {
    g_launches = 0;
    if (!theta_in || !pi_in || !theta_out || !pi_out || B <= 0 || L <= 0 || Lx <= 0) return U1_ERR_ARG;
    if (theta_in == theta_out || pi_in == pi_out) return U1_ERR_ARG;
    if (!aligned16(theta_in) || !aligned16(pi_in) || !aligned16(theta_out) || !aligned16(pi_out)) return U1_ERR_ARG;
    StepArgs a{};
    a.th = theta_in; a.pi = pi_in; a.tho = theta_out; a.pio = pi_out;
    a.Lx = Lx; a.L = L; a.beta = beta; a.drift = drift; a.kick = kick;
    return launch_step(a, B, false, (cudaStream_t)stream);
}

int u1_leapfrog_steps_rect(double* thetaA, double* piA, double* thetaB, double* piB,
                           int B, int Lx, int L, double beta, double first_drift, double dt, int n,
                           int* result_in_B, u1_stream_t stream) {
    g_launches = 0;
    if (!thetaA || !piA || !thetaB || !piB || !result_in_B || B <= 0 || L <= 0 || Lx <= 0 || n < 0) return U1_ERR_ARG;
    if (thetaA == thetaB || piA == piB) return U1_ERR_ARG;
    if (!aligned16(thetaA) || !aligned16(piA) || !aligned16(thetaB) || !aligned16(piB)) return U1_ERR_ARG;
    const double *ct = thetaA, *cp = piA;
    if (n > 0)
        U1_TRY(leapfrog_steps(thetaA, piA, thetaA, piA, thetaB, piB, B, Lx, L, beta, first_drift, dt, n,
                              (cudaStream_t)stream, &ct, &cp));
    *result_in_B = (ct == thetaB) ? 1 : 0;
    return U1_OK;
}

size_t u1_slab_reduce_workspace_bytes(int rows, int L) {
    if (rows <= 0 || L <= 0) return 0;
    const size_t ds = (size_t)((L + kSumThreads - 1) / kSumThreads) * ((rows + kDsRows - 1) / kDsRows) * 3 * sizeof(double);
    const size_t mo = (size_t)ew_grid((size_t)rows * L) * sizeof(double);
    return kTicketBytes + (ds > mo ? ds : mo);
}

int u1_slab_drift_sums_ext(const double* theta_own, const double* pi_own, size_t plane_stride, double* theta_out_own,
                           double a, double* sums, int Lx, int L, void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!theta_own || !pi_own || !theta_out_own || !sums || Lx <= 0 || L <= 0 || plane_stride < (size_t)(Lx + 1) * L)
        return U1_ERR_ARG;
    if (theta_out_own == theta_own) return U1_ERR_ARG;          // neighbours' pre-drift values are still being read
    const dim3 grid((L + kSumThreads - 1) / kSumThreads, (Lx + kDsRows - 1) / kDsRows);
    U1_TRY(check_ws(ws, ws_bytes, kTicketBytes + (size_t)grid.x * grid.y * 3 * sizeof(double)));
    k_drift_sums<<<grid, kSumThreads, 0, (cudaStream_t)stream>>>(theta_own, pi_own, theta_out_own, plane_stride, Lx, L, a,
                                                                 (double*)((char*)ws + kTicketBytes), (unsigned int*)ws, sums);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_leapfrog_steps_rect3(const double* theta_in, const double* pi_in, double* thetaA, double* piA, double* thetaB,
                            double* piB, int B, int Lx, int L, double beta, double first_drift, double dt, int n,
                            int* result_in_B, u1_stream_t stream) {
    g_launches = 0;
    if (!theta_in || !pi_in || !thetaA || !piA || !thetaB || !piB || !result_in_B || B <= 0 || L <= 0 || Lx <= 0 || n < 1)
        return U1_ERR_ARG;
    if (thetaA == thetaB || piA == piB || theta_in == thetaA || theta_in == thetaB || pi_in == piA || pi_in == piB) return U1_ERR_ARG;
    if (!aligned16(theta_in) || !aligned16(pi_in) || !aligned16(thetaA) || !aligned16(piA) || !aligned16(thetaB) || !aligned16(piB))
        return U1_ERR_ARG;
    const double *ct = nullptr, *cp = nullptr;
    U1_TRY(leapfrog_steps(theta_in, pi_in, thetaA, piA, thetaB, piB, B, Lx, L, beta, first_drift, dt, n,
                          (cudaStream_t)stream, &ct, &cp));
    *result_in_B = (ct == thetaB) ? 1 : 0;
    return U1_OK;
}

int u1_leapfrog_multi_tile_rows(int Lx) { return Lx > 0 ? (Lx + kMsOut - 1) / kMsOut : 0; }

int u1_leapfrog_multi_rows(const double* theta_in, const double* pi_in, double* theta_out, double* pi_out,
                           int Lx, int L, double beta, double first_drift, double dt, int ksteps,
                           int tile_row0, int tile_rows, u1_stream_t stream) {
    g_launches = 0;
    if (!theta_in || !pi_in || !theta_out || !pi_out || Lx <= 0 || L <= 0) return U1_ERR_ARG;
    if (theta_in == theta_out || pi_in == pi_out || ksteps < 1 || ksteps > kMsHalo) return U1_ERR_ARG;
    if (!aligned16(theta_in) || !aligned16(pi_in) || !aligned16(theta_out) || !aligned16(pi_out)) return U1_ERR_ARG;
    if (!multi_supported(Lx, L)) return U1_ERR_UNSUPPORTED;
    return launch_multi(theta_in, pi_in, theta_out, pi_out, 1, Lx, L, beta, first_drift, dt, ksteps, (cudaStream_t)stream,
                        tile_row0, tile_rows);
}

int u1_leapfrog_multi_rows2(const double* theta_in, const double* pi_in, double* theta_out, double* pi_out,
                            int Lx, int L, double beta, double first_drift, double dt, int ksteps,
                            int tile_row0, int tile_rows, int tile_row1, int tile_rows1, u1_stream_t stream) {
    g_launches = 0;
    if (!theta_in || !pi_in || !theta_out || !pi_out || Lx <= 0 || L <= 0) return U1_ERR_ARG;
    if (theta_in == theta_out || pi_in == pi_out || ksteps < 1 || ksteps > kMsHalo) return U1_ERR_ARG;
    if (!aligned16(theta_in) || !aligned16(pi_in) || !aligned16(theta_out) || !aligned16(pi_out)) return U1_ERR_ARG;
    if (!multi_supported(Lx, L)) return U1_ERR_UNSUPPORTED;
    return launch_multi(theta_in, pi_in, theta_out, pi_out, 1, Lx, L, beta, first_drift, dt, ksteps, (cudaStream_t)stream,
                        tile_row0, tile_rows, tile_row1, tile_rows1);
}

int u1_momenta_rows_dev(double* pi, uint64_t seed, uint64_t chain, uint64_t traj, const uint64_t* traj_dev,
                        long long x_first, int Lx, int L, int Lx_global, u1_stream_t stream) {
    g_launches = 0;
    if (!pi || Lx <= 0 || L <= 0 || Lx_global <= 0) return U1_ERR_ARG;
    k_momenta_rows<false><<<ew_grid((size_t)Lx * L), 256, 0, (cudaStream_t)stream>>>(pi, seed, chain, traj, traj_dev, x_first, Lx, L,
                                                                                     Lx_global, 0, 0, nullptr, nullptr, nullptr);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}
int u1_momenta_rows_pi2_dev(double* pi, uint64_t seed, uint64_t chain, uint64_t traj, const uint64_t* traj_dev,
                            long long x_first, int Lx, int L, int Lx_global, int own_first, int own_rows, double* pi2_out,
                            void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!pi || !pi2_out || Lx <= 0 || L <= 0 || Lx_global <= 0 || own_first < 0 || own_rows < 0 || own_first + own_rows > Lx)
        return U1_ERR_ARG;
    const int grid = ew_grid((size_t)Lx * L);
    U1_TRY(check_ws(ws, ws_bytes, kTicketBytes + (size_t)grid * sizeof(double)));
    k_momenta_rows<true><<<grid, 256, 0, (cudaStream_t)stream>>>(pi, seed, chain, traj, traj_dev, x_first, Lx, L, Lx_global, own_first,
                                                                 own_rows, (double*)((char*)ws + kTicketBytes), (unsigned int*)ws,
                                                                 pi2_out);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}
int u1_momenta_rows(double* pi, uint64_t seed, uint64_t chain, uint64_t traj,
                    long long x_first, int Lx, int L, int Lx_global, u1_stream_t stream) {
    return u1_momenta_rows_dev(pi, seed, chain, traj, nullptr, x_first, Lx, L, Lx_global, stream);
}

// n_steps fused steps ping-ponging between (thA,piA) and (thB,piB); returns where the result is.
static int leapfrog_stream(const double* th_in, const double* pi_in, PlainWs& w, int B, int L,
                           double beta, double dt, int n_steps, cudaStream_t st,
                           const double** th_res, const double** pi_res) {
    return leapfrog_steps(th_in, pi_in, w.thA, w.piA, w.thB, w.piB, B, L, L, beta, 0.5 * dt, dt, n_steps, st,
                          th_res, pi_res);
}

int u1_leapfrog(const double* theta_in, const double* pi_in, double* theta_out, double* pi_out,
                int B, int L, double beta, double dt, int n_steps,
                void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!theta_in || !pi_in || !theta_out || !pi_out || B <= 0 || L <= 0 || n_steps < 1) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    if (!aligned16(theta_in) || !aligned16(pi_in)) return U1_ERR_ARG;
    U1_TRY(check_ws(ws, ws_bytes, u1_workspace_bytes(B, L)));
    PlainWs w = carve_plain(ws, B, L);
    cudaStream_t st = (cudaStream_t)stream;
    const double *ct, *cp;
    U1_TRY(leapfrog_stream(theta_in, pi_in, w, B, L, beta, dt, n_steps, st, &ct, &cp));
    const size_t n = (size_t)B * 2 * L * L;
    k_final_drift<<<ew_grid(n), 256, 0, st>>>(ct, cp, theta_out, 0.5 * dt, n);
    count_launch();
    U1_CHECK_LAUNCH();
    if (pi_out != cp) {
        cudaError_t e = cudaMemcpyAsync(pi_out, cp, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
    }
    return U1_OK;
}

int u1_trajectory(double* theta, const double* pi_inject, const double* u_inject,
                  uint64_t seed, uint64_t chain0, uint64_t traj0, int n_traj,
                  int B, int L, double beta, double dt, int n_steps,
                  double* out_dH, int32_t* out_accept, double* out_H, double* out_plaq,
                  double* out_topo, int path, void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!theta || B <= 0 || L <= 0 || n_traj < 1 || n_steps < 1) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    // both kernel paths move fields as double2 / TMA bulk copies: 16-byte aligned chain bases are required
    // (L even makes every chain base aligned once the first one is)
    if (!aligned16(theta) || (pi_inject && !aligned16(pi_inject))) return U1_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool resident = (path == 2) || (path == 0 && resident_supported(L));
    if (resident) {
        if (!resident_supported(L)) return U1_ERR_UNSUPPORTED;
        TrajArgs a{};
        a.theta = theta; a.pi_inject = pi_inject; a.u_inject = u_inject;
        a.seed = seed; a.chain0 = chain0; a.traj0 = traj0;
        a.n_traj = n_traj; a.B = B; a.n_steps = n_steps; a.beta = beta; a.dt = dt;
        a.out_dH = out_dH; a.out_acc = out_accept; a.out_H = out_H; a.out_plaq = out_plaq; a.out_topo = out_topo;
        return launch_resident(a, L, st);
    }
    U1_TRY(check_ws(ws, ws_bytes, u1_workspace_bytes(B, L)));
    PlainWs w = carve_plain(ws, B, L);
    const size_t per = (size_t)2 * L * L, n = (size_t)B * per;
    const double inv_vol = 1.0 / ((double)L * L);
    for (int t = 0; t < n_traj; ++t) {
        const uint64_t traj = traj0 + (uint64_t)t;
        const double* pi0;
        if (pi_inject) pi0 = pi_inject + (size_t)t * n;
        else {
            k_momenta<<<ew_grid((size_t)B * L * L), 256, 0, st>>>(w.piB, seed, chain0, traj, B, L);
            count_launch();
            U1_CHECK_LAUNCH();
            pi0 = w.piB;   // consumed by step 0 (which writes thA/piA) before piB is overwritten
        }
        U1_TRY(launch_sums(theta, nullptr, pi0, w.s_old, w.partial, B, L, L, st));
        const double *ct, *cp;
        U1_TRY(leapfrog_stream(theta, pi0, w, B, L, beta, dt, n_steps, st, &ct, &cp));
        double* fin = (ct == w.thA) ? w.thB : w.thA;
        k_final_drift<<<ew_grid(n), 256, 0, st>>>(ct, cp, fin, 0.5 * dt, n);
        count_launch();
        U1_CHECK_LAUNCH();
        U1_TRY(launch_sums(fin, nullptr, cp, w.s_new, w.partial, B, L, L, st));
        const size_t o = (size_t)t * B;
        k_metropolis<<<(B + 127) / 128, 128, 0, st>>>(
            w.s_old, w.s_new, u_inject ? u_inject + o : nullptr, seed, chain0, traj, nullptr, B, beta, inv_vol,
            out_dH ? out_dH + o : nullptr, out_accept ? out_accept + o : nullptr,
            out_H ? out_H + o : nullptr, out_plaq ? out_plaq + o : nullptr,
            out_topo ? out_topo + o : nullptr, w.acc);
        count_launch();
        U1_CHECK_LAUNCH();
        k_select<<<ew_grid(n), 256, 0, st>>>(theta, fin, w.acc, per, n);
        count_launch();
        U1_CHECK_LAUNCH();
    }
    return U1_OK;
}

int u1_slab_leapfrog_step(const double* theta_in, const double* pi_in,
                          const double* theta_up, const double* pi_up,
                          const double* theta_dn, const double* pi_dn,
                          double* theta_out, double* pi_out,
                          int Lx, int L, double beta, double drift, double kick,
                          u1_stream_t stream) {
    g_launches = 0;
    if (!theta_in || !pi_in || !theta_up || !pi_up || !theta_dn || !pi_dn || !theta_out || !pi_out ||
        Lx <= 0 || L <= 0)
        return U1_ERR_ARG;
    if (theta_in == theta_out || pi_in == pi_out) return U1_ERR_ARG;
    if (!aligned16(theta_in) || !aligned16(pi_in) || !aligned16(theta_out) || !aligned16(pi_out)) return U1_ERR_ARG;
    StepArgs a{};
    a.th = theta_in; a.pi = pi_in; a.tho = theta_out; a.pio = pi_out;
    a.th_up = theta_up; a.pi_up = pi_up; a.th_dn = theta_dn; a.pi_dn = pi_dn;
    a.Lx = Lx; a.L = L; a.beta = beta; a.drift = drift; a.kick = kick;
    return launch_step(a, 1, true, (cudaStream_t)stream);
}

int u1_drift_wrap(const double* theta, const double* pi, double* out, double a, size_t n,
                  u1_stream_t stream) {
    g_launches = 0;
    if (!theta || !pi || !out) return U1_ERR_ARG;
    if (n == 0) return U1_OK;
    k_final_drift<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(theta, pi, out, a, n);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_metropolis_sums_dev(const double* sums_old, const double* sums_new, const double* u_inject,
                           uint64_t seed, uint64_t chain0, uint64_t traj, const uint64_t* traj_dev, int B,
                           double beta, double volume, double* out_dH, int32_t* out_accept, double* out_H,
                           double* out_plaq, double* out_topo, u1_stream_t stream) {
    g_launches = 0;
    if (!sums_old || !sums_new || !out_accept || B <= 0 || !(volume > 0)) return U1_ERR_ARG;
    k_metropolis<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        sums_old, sums_new, u_inject, seed, chain0, traj, traj_dev, B, beta, 1.0 / volume, out_dH, nullptr, out_H,
        out_plaq, out_topo, out_accept);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}
int u1_metropolis_sums(const double* sums_old, const double* sums_new, const double* u_inject,
                       uint64_t seed, uint64_t chain0, uint64_t traj, int B, double beta, double volume,
                       double* out_dH, int32_t* out_accept, double* out_H, double* out_plaq,
                       double* out_topo, u1_stream_t stream) {
    return u1_metropolis_sums_dev(sums_old, sums_new, u_inject, seed, chain0, traj, nullptr, B, beta, volume,
                                  out_dH, out_accept, out_H, out_plaq, out_topo, stream);
}

int u1_select_accepted(double* theta, const double* theta_new, const int32_t* accept,
                       size_t per_chain, int B, u1_stream_t stream) {
    g_launches = 0;
    if (!theta || !theta_new || !accept || B <= 0 || per_chain == 0) return U1_ERR_ARG;
    const size_t n = per_chain * (size_t)B;
    k_select<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(theta, theta_new, accept, per_chain, n);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1_slab_sums(const double* theta, const double* theta_dn, const double* pi, double* sums,
                 int Lx, int L, void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!theta || !theta_dn || !sums || Lx <= 0 || L <= 0) return U1_ERR_ARG;
    const size_t need = (size_t)sum_blocks(1, Lx) * 3 * sizeof(double);
    U1_TRY(check_ws(ws, ws_bytes, need));
    return launch_sums(theta, theta_dn, pi, sums, (double*)ws, 1, Lx, L, (cudaStream_t)stream);
}

int u1_plain_preload_slab_kernels(void) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, k_leapfrog_multi<32>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_leapfrog_multi<64>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_leapfrog_step<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_leapfrog_step<true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_partial_sums);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_finish_sums);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_momenta_rows<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_momenta_rows<true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_drift_sums);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_metropolis);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_select);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k_final_drift);
    return e == cudaSuccess ? U1_OK : (int)e;
}

int u1_slab_sums_ext(const double* theta_own, const double* pi_own, size_t plane_stride, double* sums,
                     int Lx, int L, void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    if (!theta_own || !sums || Lx <= 0 || L <= 0 || plane_stride < (size_t)(Lx + 1) * L) return U1_ERR_ARG;
    const size_t need = (size_t)sum_blocks(1, Lx) * 3 * sizeof(double);
    U1_TRY(check_ws(ws, ws_bytes, need));
    return launch_sums(theta_own, nullptr, pi_own, sums, (double*)ws, 1, Lx, L, (cudaStream_t)stream, plane_stride, 1);
}

}  // extern "C"
