// Standalone microbenchmark of fp64 3x3 circular conv (12->12) inner-loop variants on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/conv_variants tools/microbench/conv_variants.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cmath>

template <int CIN, int COUT> struct alignas(16) ConvP { double w[CIN * 9 * COUT]; double b[COUT]; };

// ROWS vertically adjacent sites per thread, weights read as double2 (WV=2) or double (WV=1)
template <int ROWS, int WV, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_conv(const __grid_constant__ ConvP<12, 12> p, const double* __restrict__ in, double* __restrict__ out, int ngroups, int L) {
    constexpr int CIN = 12, COUT = 12;
    const int s = blockIdx.x * THREADS + threadIdx.x;
    if (s >= ngroups) return;
    const int V = L * L, GV = V / ROWS;
    const int b = s / GV, rp = s - b * GV;
    const int xp = rp / L, y = rp - xp * L, x = ROWS * xp;
    int xr[ROWS + 2];
#pragma unroll
    for (int i = 0; i < ROWS + 2; ++i) { int xx = x + i - 1; xx = xx < 0 ? xx + L : (xx >= L ? xx - L : xx); xr[i] = xx * L; }
    const int yc[3] = {(y == 0 ? L - 1 : y - 1), y, (y == L - 1 ? 0 : y + 1)};
    const double* ib = in + (size_t)b * CIN * V;
    double acc[ROWS][COUT];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[r][o] = 0.0;
#pragma unroll 1
    for (int c = 0; c < CIN; ++c) {
        const double* pc = ib + (size_t)c * V;
        double v[ROWS + 2][3];
#pragma unroll
        for (int i = 0; i < ROWS + 2; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) v[i][j] = pc[xr[i] + yc[j]];
        if (WV == 2) {
            const double2* w2 = reinterpret_cast<const double2*>(p.w) + c * 9 * COUT / 2;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int o = 0; o < COUT; o += 2) {
                        const double2 wv = w2[((i * 3 + j) * COUT + o) / 2];
#pragma unroll
                        for (int r = 0; r < ROWS; ++r) {
                            acc[r][o] = fma(wv.x, v[i + r][j], acc[r][o]);
                            acc[r][o + 1] = fma(wv.y, v[i + r][j], acc[r][o + 1]);
                        }
                    }
        } else {
            const double* w = p.w + c * 9 * COUT;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int o = 0; o < COUT; ++o) {
                        const double wv = w[(i * 3 + j) * COUT + o];
#pragma unroll
                        for (int r = 0; r < ROWS; ++r) acc[r][o] = fma(wv, v[i + r][j], acc[r][o]);
                    }
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) out[(size_t)b * COUT * V + (size_t)o * V + (size_t)(x + r) * L + y] = acc[r][o];
}

// experiments: channel-loop unroll factor U; FAKE=1 no constant-bank weights, FAKE=2 no global loads in the loop
template <int ROWS, int U, int FAKE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_conv_x(const __grid_constant__ ConvP<12, 12> p, const double* __restrict__ in, double* __restrict__ out, int ngroups, int L) {
    constexpr int CIN = 12, COUT = 12;
    const int s = blockIdx.x * THREADS + threadIdx.x;
    if (s >= ngroups) return;
    const int V = L * L, GV = V / ROWS;
    const int b = s / GV, rp = s - b * GV;
    const int xp = rp / L, y = rp - xp * L, x = ROWS * xp;
    int xr[ROWS + 2];
#pragma unroll
    for (int i = 0; i < ROWS + 2; ++i) { int xx = x + i - 1; xx = xx < 0 ? xx + L : (xx >= L ? xx - L : xx); xr[i] = xx * L; }
    const int yc[3] = {(y == 0 ? L - 1 : y - 1), y, (y == L - 1 ? 0 : y + 1)};
    const double* ib = in + (size_t)b * CIN * V;
    double acc[ROWS][COUT];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[r][o] = 0.0;
    double fw = in[s & 1023], fv = in[(s + 7) & 1023];
#pragma unroll U
    for (int c = 0; c < CIN; ++c) {
        const double* pc = ib + (size_t)c * V;
        double v[ROWS + 2][3];
#pragma unroll
        for (int i = 0; i < ROWS + 2; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) v[i][j] = (FAKE == 2 || FAKE >= 3) ? fv + (double)(i * 3 + j + c) : pc[xr[i] + yc[j]];
        const double* w = p.w + c * 9 * COUT;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int o = 0; o < COUT; ++o) {
                    const double wv = (FAKE == 1) ? fw : (FAKE >= 3 ? p.w[o + (FAKE == 4 ? 12 * j : 0)] : w[(i * 3 + j) * COUT + o]);
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) acc[r][o] = fma(wv, v[i + r][j], acc[r][o]);
                }
        if (FAKE == 1) fw += 1e-9;
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) out[(size_t)b * COUT * V + (size_t)o * V + (size_t)(x + r) * L + y] = acc[r][o];
}

// TR x TC sites per thread (TC even, columns contiguous -> double2 loads), output channels processed in
// NH passes of COUT/NH channels each (fewer live accumulators -> more sites per weight load).
template <int TR, int TC, int NH, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_conv_tile(const __grid_constant__ ConvP<12, 12> p, const double* __restrict__ in, double* __restrict__ out, int ngroups, int L) {
    constexpr int CIN = 12, COUT = 12, CH = COUT / NH;
    const int s = blockIdx.x * THREADS + threadIdx.x;
    if (s >= ngroups) return;
    const int V = L * L, GW = L / TC, GV = (L / TR) * GW;
    const int b = s / GV, rp = s - b * GV;
    const int xt = rp / GW, yt = rp - xt * GW;
    const int x = xt * TR, y = yt * TC;
    int xr[TR + 2];
#pragma unroll
    for (int i = 0; i < TR + 2; ++i) { int xx = x + i - 1; xx = xx < 0 ? xx + L : (xx >= L ? xx - L : xx); xr[i] = xx * L; }
    const int ym = (y == 0 ? L - 1 : y - 1), yp = (y + TC >= L ? 0 : y + TC);
    const double* ib = in + (size_t)b * CIN * V;
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
        double acc[TR][TC][CH];
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int q = 0; q < TC; ++q)
#pragma unroll
                for (int o = 0; o < CH; ++o) acc[r][q][o] = 0.0;
#pragma unroll 1
        for (int c = 0; c < CIN; ++c) {
            const double* pc = ib + (size_t)c * V;
            double v[TR + 2][TC + 2];
#pragma unroll
            for (int i = 0; i < TR + 2; ++i) {
                v[i][0] = pc[xr[i] + ym];
#pragma unroll
                for (int q = 0; q < TC; q += 2) {
                    const double2 t = *reinterpret_cast<const double2*>(pc + xr[i] + y + q);
                    v[i][1 + q] = t.x; v[i][2 + q] = t.y;
                }
                v[i][TC + 1] = pc[xr[i] + yp];
            }
            const double* w = p.w + c * 9 * COUT + h * CH;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int o = 0; o < CH; ++o) {
                        const double wv = w[(i * 3 + j) * COUT + o];
#pragma unroll
                        for (int r = 0; r < TR; ++r)
#pragma unroll
                            for (int q = 0; q < TC; ++q) acc[r][q][o] = fma(wv, v[i + r][j + q], acc[r][q][o]);
                    }
        }
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int o = 0; o < CH; ++o)
#pragma unroll
                for (int q = 0; q < TC; q += 2)
                    *reinterpret_cast<double2*>(out + (size_t)b * COUT * V + (size_t)(h * CH + o) * V + (size_t)(x + r) * L + y + q) =
                        make_double2(acc[r][q][o], acc[r][q + 1][o]);
    }
}

// input tile staged in shared memory with cp.async (STAGES-deep ring over channels); block = THREADS
// threads = (THREADS/L) groups of ROWS full rows (L in {16,32,64,128}); weights from the constant bank.
template <int LL, int ROWS, int STAGES, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_conv_cpasync(const __grid_constant__ ConvP<12, 12> p, const double* __restrict__ in, double* __restrict__ out, int ngroups, int L) {
    constexpr int CIN = 12, COUT = 12;
    constexpr int TP = THREADS / LL;                  // row groups per block
    constexpr int TROWS = ROWS * TP + 2;              // tile rows incl. halo
    constexpr int TILE = TROWS * LL;                  // doubles per channel tile
    __shared__ __align__(16) double tile[STAGES][TILE];
    const int V = LL * LL;
    const int groups_per_chain = V / ROWS;
    const int s0 = blockIdx.x * THREADS;
    const int b = s0 / groups_per_chain, rp0 = s0 - b * groups_per_chain;
    const int xg0 = rp0 / LL;                         // first row group
    const int tid = threadIdx.x;
    const int tp = tid / LL, y = tid - tp * LL;
    const int x = ROWS * (xg0 + tp);
    const double* ib = in + (size_t)b * CIN * V;
    auto issue = [&](int c, int stage) {
        const double* pc = ib + (size_t)c * V;
        constexpr int CHUNKS = TILE / 2;
        for (int q = tid; q < CHUNKS; q += THREADS) {
            const int r = (2 * q) / LL, col = 2 * q - r * LL;
            int xx = ROWS * xg0 - 1 + r; xx = xx < 0 ? xx + LL : (xx >= LL ? xx - LL : xx);
            const double* src = pc + xx * LL + col;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&tile[stage][r * LL + col]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
    };
    double acc[ROWS][COUT];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[r][o] = 0.0;
#pragma unroll
    for (int st = 0; st < STAGES - 1; ++st) issue(st, st);
    const int ym = (y == 0 ? LL - 1 : y - 1), yp = (y == LL - 1 ? 0 : y + 1);
    const int r0 = ROWS * tp;
#pragma unroll 1
    for (int c = 0; c < CIN; ++c) {
        // groups committed so far: min(CIN, c + STAGES - 1); need group c complete
        if (c + STAGES - 1 <= CIN) asm volatile("cp.async.wait_group %0;" :: "n"(STAGES - 2));
        else asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        if (c + STAGES - 1 < CIN) issue(c + STAGES - 1, (c + STAGES - 1) % STAGES);
        const double* t = tile[c % STAGES];
        double v[ROWS + 2][3];
#pragma unroll
        for (int i = 0; i < ROWS + 2; ++i) { v[i][0] = t[(r0 + i) * LL + ym]; v[i][1] = t[(r0 + i) * LL + y]; v[i][2] = t[(r0 + i) * LL + yp]; }
        const double* w = p.w + c * 9 * COUT;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int o = 0; o < COUT; ++o) {
                    const double wv = w[(i * 3 + j) * COUT + o];
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) acc[r][o] = fma(wv, v[i + r][j], acc[r][o]);
                }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) out[(size_t)b * COUT * V + (size_t)o * V + (size_t)(x + r) * LL + y] = acc[r][o];
}

// ---- forward-layer variants: same inner loops plus the table-GELU epilogue of the library ----
constexpr int kGI = 136, kGD = 11, kGS = 13;
__device__ __forceinline__ void gelu_tab_eval(const double* __restrict__ tab, double z, double& a, double& d) {
    const double zc = fmin(fmax(z, -8.5), 8.5 - 1e-9);
    const int i = __double2int_rd((zc + 8.5) * 8.0);
    const double t = zc - (((double)i + 0.5) * 0.125 - 8.5);
    const double* q = tab + i * kGS;
    double p = q[kGD], dp = 0.0;
#pragma unroll
    for (int k = kGD - 1; k >= 0; --k) { dp = fma(dp, t, p); p = fma(p, t, q[k]); }
    a = z * p; d = fma(z, dp, p);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k_fwd_direct(const __grid_constant__ ConvP<12, 12> p, const double* __restrict__ in, double* __restrict__ out, double* __restrict__ dout,
             const double* __restrict__ gtab, int npairs, int L) {
    constexpr int CIN = 12, COUT = 12;
    extern __shared__ double s_tab[];
    for (int i = threadIdx.x; i < kGI * kGS; i += 128) s_tab[i] = gtab[i];
    __syncthreads();
    const int s = blockIdx.x * 128 + threadIdx.x;
    if (s >= npairs) return;
    const int V = L * L, HV = V >> 1;
    const int b = s / HV, rp = s - b * HV;
    const int xp = rp / L, y = rp - xp * L, x = 2 * xp;
    const int xr[4] = {(x == 0 ? L - 1 : x - 1) * L, x * L, (x + 1) * L, (x + 2 >= L ? 0 : x + 2) * L};
    const int yc[3] = {(y == 0 ? L - 1 : y - 1), y, (y == L - 1 ? 0 : y + 1)};
    const double* ib = in + (size_t)b * CIN * V;
    double a0[COUT], a1[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) { a0[o] = 0.0; a1[o] = 0.0; }
#pragma unroll 1
    for (int c = 0; c < CIN; ++c) {
        const double* pc = ib + (size_t)c * V;
        double v[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) v[i][j] = pc[xr[i] + yc[j]];
        const double* w = p.w + c * 9 * COUT;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int o = 0; o < COUT; ++o) { const double wv = w[(i * 3 + j) * COUT + o]; a0[o] = fma(wv, v[i][j], a0[o]); a1[o] = fma(wv, v[i + 1][j], a1[o]); }
    }
    const size_t ob = (size_t)b * COUT * V + (size_t)x * L + y;
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
        const size_t i0 = ob + (size_t)o * V, i1 = i0 + L;
        double g0, d0, g1, d1;
        gelu_tab_eval(s_tab, a0[o] + p.b[o], g0, d0);
        gelu_tab_eval(s_tab, a1[o] + p.b[o], g1, d1);
        out[i0] = g0; out[i1] = g1; dout[i0] = d0; dout[i1] = d1;
    }
}

template <int LL, int ROWS, int NH, int MINB>      // cp.async tile, ROWS rows per thread, output channels in NH passes
__global__ void __launch_bounds__(128, MINB)
k_fwd_tile(const __grid_constant__ ConvP<12, 12> p, const double* __restrict__ in, double* __restrict__ out, double* __restrict__ dout,
           const double* __restrict__ gtab) {
    constexpr int CIN = 12, COUT = 12, CH = COUT / NH, STAGES = 3, THREADS = 128;
    constexpr int TP = THREADS / LL, TROWS = ROWS * TP + 2, TILE = TROWS * LL, V = LL * LL;
    extern __shared__ __align__(16) double s_dyn[];
    double* tile = s_dyn;
    double* s_tab = s_dyn + STAGES * TILE;
    for (int i = threadIdx.x; i < kGI * kGS; i += THREADS) s_tab[i] = gtab[i];
    constexpr int groups_per_chain = V / ROWS;
    const int s0 = blockIdx.x * THREADS;
    const int b = s0 / groups_per_chain, rp0 = s0 - b * groups_per_chain;
    const int xg0 = rp0 / LL, tid = threadIdx.x, tp = tid / LL, y = tid - tp * LL, x = ROWS * (xg0 + tp);
    const double* ib = in + (size_t)b * CIN * V;
    auto issue = [&](int c, int stage) {
        const double* pc = ib + (size_t)c * V;
        for (int q = tid; q < TILE / 2; q += THREADS) {
            const int r = (2 * q) / LL, col = 2 * q - r * LL;
            int xx = ROWS * xg0 - 1 + r; xx = xx < 0 ? xx + LL : (xx >= LL ? xx - LL : xx);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + stage * TILE + r * LL + col);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(pc + xx * LL + col));
        }
        asm volatile("cp.async.commit_group;");
    };
    const int ym = (y == 0 ? LL - 1 : y - 1), yp = (y == LL - 1 ? 0 : y + 1), r0 = ROWS * tp;
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
        double acc[ROWS][CH];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int o = 0; o < CH; ++o) acc[r][o] = 0.0;
        __syncthreads();                                  // ring buffers free again (second pass)
        issue(0, 0); issue(1, 1);
#pragma unroll 1
        for (int c = 0; c < CIN; ++c) {
            if (c + 2 <= CIN) asm volatile("cp.async.wait_group 1;"); else asm volatile("cp.async.wait_group 0;");
            __syncthreads();
            if (c + 2 < CIN) issue(c + 2, (c + 2) % STAGES);
            const double* t = tile + (c % STAGES) * TILE;
            double v[ROWS + 2][3];
#pragma unroll
            for (int i = 0; i < ROWS + 2; ++i) { v[i][0] = t[(r0 + i) * LL + ym]; v[i][1] = t[(r0 + i) * LL + y]; v[i][2] = t[(r0 + i) * LL + yp]; }
            const double* w = p.w + c * 9 * COUT + h * CH;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int o = 0; o < CH; ++o) {
                        const double wv = w[(i * 3 + j) * COUT + o];
#pragma unroll
                        for (int r = 0; r < ROWS; ++r) acc[r][o] = fma(wv, v[i + r][j], acc[r][o]);
                    }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int o = 0; o < CH; ++o) {
                const size_t idx = (size_t)b * COUT * V + (size_t)(h * CH + o) * V + (size_t)(x + r) * LL + y;
                double g, d;
                gelu_tab_eval(s_tab, acc[r][o] + p.b[h * CH + o], g, d);
                out[idx] = g; dout[idx] = d;
            }
    }
}

// software-pipelined: the window of channel c+1 is loaded while channel c is being accumulated
template <int ROWS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_conv_pf(const __grid_constant__ ConvP<12, 12> p, const double* __restrict__ in, double* __restrict__ out, int ngroups, int L) {
    constexpr int CIN = 12, COUT = 12;
    const int s = blockIdx.x * THREADS + threadIdx.x;
    if (s >= ngroups) return;
    const int V = L * L, GV = V / ROWS;
    const int b = s / GV, rp = s - b * GV;
    const int xp = rp / L, y = rp - xp * L, x = ROWS * xp;
    int off[ROWS + 2][3];
    {
        const int yc[3] = {(y == 0 ? L - 1 : y - 1), y, (y == L - 1 ? 0 : y + 1)};
#pragma unroll
        for (int i = 0; i < ROWS + 2; ++i) {
            int xx = x + i - 1; xx = xx < 0 ? xx + L : (xx >= L ? xx - L : xx);
#pragma unroll
            for (int j = 0; j < 3; ++j) off[i][j] = xx * L + yc[j];
        }
    }
    const double* ib = in + (size_t)b * CIN * V;
    double acc[ROWS][COUT];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[r][o] = 0.0;
    double va[ROWS + 2][3], vb[ROWS + 2][3];
    auto load = [&](double (&v)[ROWS + 2][3], int c) {
        const double* pc = ib + (size_t)c * V;
#pragma unroll
        for (int i = 0; i < ROWS + 2; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) v[i][j] = pc[off[i][j]];
    };
    auto mac = [&](const double (&v)[ROWS + 2][3], int c) {
        const double* w = p.w + c * 9 * COUT;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int o = 0; o < COUT; ++o) {
                    const double wv = w[(i * 3 + j) * COUT + o];
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) acc[r][o] = fma(wv, v[i + r][j], acc[r][o]);
                }
    };
    load(va, 0);
#pragma unroll 1
    for (int c = 0; c < CIN; c += 2) {
        load(vb, c + 1);
        mac(va, c);
        load(va, (c + 2 < CIN) ? c + 2 : CIN - 1);      // unconditional (clamped) so it is not sunk below the MACs
        mac(vb, c + 1);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) out[(size_t)b * COUT * V + (size_t)o * V + (size_t)(x + r) * L + y] = acc[r][o];
}

// weights staged in shared memory (LDS.128 broadcast), ROWS sites per thread
template <int ROWS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_conv_smem(const __grid_constant__ ConvP<12, 12> p, const double* __restrict__ in, double* __restrict__ out, int ngroups, int L) {
    constexpr int CIN = 12, COUT = 12;
    __shared__ double2 sw[CIN * 9 * COUT / 2];
    for (int i = threadIdx.x; i < CIN * 9 * COUT / 2; i += THREADS) sw[i] = make_double2(p.w[2 * i], p.w[2 * i + 1]);
    __syncthreads();
    const int s = blockIdx.x * THREADS + threadIdx.x;
    if (s >= ngroups) return;
    const int V = L * L, GV = V / ROWS;
    const int b = s / GV, rp = s - b * GV;
    const int xp = rp / L, y = rp - xp * L, x = ROWS * xp;
    int xr[ROWS + 2];
#pragma unroll
    for (int i = 0; i < ROWS + 2; ++i) { int xx = x + i - 1; xx = xx < 0 ? xx + L : (xx >= L ? xx - L : xx); xr[i] = xx * L; }
    const int yc[3] = {(y == 0 ? L - 1 : y - 1), y, (y == L - 1 ? 0 : y + 1)};
    const double* ib = in + (size_t)b * CIN * V;
    double acc[ROWS][COUT];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[r][o] = 0.0;
#pragma unroll 1
    for (int c = 0; c < CIN; ++c) {
        const double* pc = ib + (size_t)c * V;
        double v[ROWS + 2][3];
#pragma unroll
        for (int i = 0; i < ROWS + 2; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) v[i][j] = pc[xr[i] + yc[j]];
        const double2* w2 = sw + c * 9 * COUT / 2;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int o = 0; o < COUT; o += 2) {
                    const double2 wv = w2[((i * 3 + j) * COUT + o) / 2];
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
                        acc[r][o] = fma(wv.x, v[i + r][j], acc[r][o]);
                        acc[r][o + 1] = fma(wv.y, v[i + r][j], acc[r][o + 1]);
                    }
                }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) out[(size_t)b * COUT * V + (size_t)o * V + (size_t)(x + r) * L + y] = acc[r][o];
}

template <typename F>
float time_it(F f, int reps = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char** argv) {
    const int L = 64, B = argc > 1 ? atoi(argv[1]) : 166, V = L * L;
    const size_t n = (size_t)B * 12 * V;
    double *in, *out, *out2;
    cudaMalloc(&in, n * 8); cudaMalloc(&out, n * 8); cudaMalloc(&out2, n * 8);
    std::vector<double> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = sin(0.001 * i);
    cudaMemcpy(in, h.data(), n * 8, cudaMemcpyHostToDevice);
    ConvP<12, 12> p;
    for (int i = 0; i < 12 * 9 * 12; ++i) p.w[i] = cos(0.37 * i) * 0.1;
    const double ideal_us = (double)B * V * 1296 / (148.0 * 64 * 1.9e9) * 1e6;
    printf("B=%d sites=%d ideal(1.9GHz)=%.1f us\n", B, B * V, ideal_us);
    auto report = [&](const char* name, float ms) { printf("%-34s %8.1f us  %5.1f%% of DFMA peak@1.9GHz\n", name, ms * 1e3, ideal_us / (ms * 1e3) * 100); };
#define RUN(ROWS, WV, T, MB) { int ng = B * V / ROWS; float ms = time_it([&] { k_conv<ROWS, WV, T, MB><<<(ng + T - 1) / T, T>>>(p, in, out, ng, L); }); \
        report("rows" #ROWS " wv" #WV " t" #T " minb" #MB, ms); }
#define RUNS(ROWS, T, MB) { int ng = B * V / ROWS; float ms = time_it([&] { k_conv_smem<ROWS, T, MB><<<(ng + T - 1) / T, T>>>(p, in, out2, ng, L); }); \
        report("SMEM rows" #ROWS " t" #T " minb" #MB, ms); }
#define RUNP(ROWS, T, MB) { int ng = B * V / ROWS; float ms = time_it([&] { k_conv_pf<ROWS, T, MB><<<(ng + T - 1) / T, T>>>(p, in, out2, ng, L); }); \
        report("PREFETCH rows" #ROWS " t" #T " minb" #MB, ms); }
#define RUNX(ROWS, U, FAKE, T, MB) { int ng = B * V / ROWS; float ms = time_it([&] { k_conv_x<ROWS, U, FAKE, T, MB><<<(ng + T - 1) / T, T>>>(p, in, out2, ng, L); }); \
        report("X rows" #ROWS " unroll" #U " fake" #FAKE " t" #T " minb" #MB, ms); }
    if (argc > 2 && atoi(argv[2]) == 8) {
        RUNP(2, 128, 1) RUNP(2, 128, 4) RUNP(2, 128, 5) RUNP(2, 64, 8) RUNP(2, 256, 2) RUNP(4, 128, 1) RUNP(4, 128, 2) RUNP(4, 128, 3) RUNP(4, 64, 4) RUNP(4, 64, 6) RUNP(1, 128, 1) RUNP(1, 128, 8)
        return 0;
    }
#define RUNT(TR, TC, NH, T, MB) { int ng = B * V / (TR * TC); float ms = time_it([&] { k_conv_tile<TR, TC, NH, T, MB><<<(ng + T - 1) / T, T>>>(p, in, out2, ng, L); }); \
        report("TILE " #TR "x" #TC " halves" #NH " t" #T " minb" #MB, ms); }
    if (argc > 2 && atoi(argv[2]) == 4) {
        double* gtab; cudaMalloc(&gtab, kGI * kGS * 8);
        { std::vector<double> t(kGI * kGS); for (int i = 0; i < kGI * kGS; ++i) t[i] = 0.01 * cos(0.1 * i); cudaMemcpy(gtab, t.data(), t.size() * 8, cudaMemcpyHostToDevice); }
        double* dout; cudaMalloc(&dout, n * 8);
        const size_t gsm = kGI * kGS * 8;
#define RUNFD(MB) { int ng = B * V / 2; float ms = time_it([&] { k_fwd_direct<MB><<<(ng + 127) / 128, 128, gsm>>>(p, in, out, dout, gtab, ng, L); }); report("FWD direct rows2 minb" #MB, ms); }
#define RUNFT(ROWS, NH, MB) { constexpr size_t sm = (3 * (ROWS * 2 + 2) * 64 + kGI * kGS) * 8; \
        cudaFuncSetAttribute(k_fwd_tile<64, ROWS, NH, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        int nb = B * V / (ROWS * 128); float ms = time_it([&] { k_fwd_tile<64, ROWS, NH, MB><<<nb, 128, sm>>>(p, in, out2, dout, gtab); }); \
        report("FWD tile rows" #ROWS " passes" #NH " minb" #MB, ms); }
        RUNFD(1) RUNFD(6) RUNFD(5)
        RUNFT(2, 1, 5) RUNFT(2, 1, 6) RUNFT(2, 1, 4) RUNFT(4, 2, 4) RUNFT(4, 2, 5) RUNFT(4, 2, 6) RUNFT(4, 2, 3) RUNFT(2, 2, 6) RUNFT(2, 2, 8)
        { int ng = B * V / 2; k_fwd_direct<1><<<(ng + 127) / 128, 128, gsm>>>(p, in, out, dout, gtab, ng, L); constexpr size_t sm = (3 * 10 * 64 + kGI * kGS) * 8; k_fwd_tile<64, 4, 2, 4><<<B * V / 512, 128, sm>>>(p, in, out2, dout, gtab); }
        std::vector<double> a(n), b2(n);
        cudaMemcpy(a.data(), out, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b2.data(), out2, n * 8, cudaMemcpyDeviceToHost);
        double md = 0; for (size_t i = 0; i < n; ++i) md = fmax(md, fabs(a[i] - b2[i]));
        printf("max diff fwd tile vs direct %.3e (%s)\n", md, cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    if (argc > 2 && atoi(argv[2]) == 5) {
        RUN(2, 2, 128, 1)
#define RUNC(ROWS, ST, T, MB) { int ng = B * V / ROWS; float ms = time_it([&] { k_conv_cpasync<64, ROWS, ST, T, MB><<<ng / T, T>>>(p, in, out2, ng, L); }); \
        report("CPASYNC rows" #ROWS " stages" #ST " t" #T " minb" #MB, ms); }
        RUNC(2, 3, 128, 6) RUNC(2, 3, 128, 5) RUNC(2, 3, 128, 7) RUNC(2, 4, 128, 6) RUNC(2, 2, 128, 6) RUNC(2, 3, 256, 3) RUNC(2, 3, 64, 12)
        RUNC(4, 3, 128, 3) RUNC(4, 3, 128, 4) RUNC(4, 3, 64, 6) RUNC(4, 3, 64, 8) RUNC(4, 4, 128, 3) RUNC(1, 3, 128, 8) RUNC(1, 3, 128, 12)
        { int ng = B * V / 2; k_conv<2, 2, 128, 1><<<(ng + 127) / 128, 128>>>(p, in, out, ng, L); k_conv_cpasync<64, 4, 3, 128, 3><<<(ng / 2) / 128, 128>>>(p, in, out2, ng / 2, L); }
        std::vector<double> a(n), b2(n);
        cudaMemcpy(a.data(), out, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b2.data(), out2, n * 8, cudaMemcpyDeviceToHost);
        double md = 0; for (size_t i = 0; i < n; ++i) md = fmax(md, fabs(a[i] - b2[i]));
        printf("max diff cpasync vs rows2 %.3e (%s)\n", md, cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    if (argc > 2 && atoi(argv[2]) == 6) {
        RUNT(2, 2, 1, 128, 1) RUNT(2, 2, 1, 128, 3) RUNT(2, 2, 1, 64, 6) RUNT(1, 4, 1, 128, 3) RUNT(1, 2, 1, 128, 1)
        RUNT(2, 4, 2, 128, 1) RUNT(2, 4, 2, 128, 2) RUNT(2, 4, 2, 128, 3) RUNT(2, 4, 2, 64, 4) RUNT(2, 4, 2, 64, 6)
        RUNT(4, 2, 2, 128, 2) RUNT(4, 2, 2, 64, 4) RUNT(2, 2, 2, 128, 4) RUNT(2, 2, 2, 128, 6) RUNT(2, 4, 3, 128, 3) RUNT(2, 4, 4, 128, 4) RUNT(4, 4, 4, 128, 2) RUNT(4, 4, 4, 64, 4)
        { int ng = B * V / 2; k_conv<2, 2, 128, 1><<<(ng + 127) / 128, 128>>>(p, in, out, ng, L); ng = B * V / 8; k_conv_tile<2, 4, 2, 128, 2><<<(ng + 127) / 128, 128>>>(p, in, out2, ng, L); }
        std::vector<double> a(n), b2(n);
        cudaMemcpy(a.data(), out, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b2.data(), out2, n * 8, cudaMemcpyDeviceToHost);
        double md = 0; for (size_t i = 0; i < n; ++i) md = fmax(md, fabs(a[i] - b2[i]));
        printf("max diff tile vs rows2 %.3e (%s)\n", md, cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    if (argc > 2 && atoi(argv[2]) == 7) {
        RUNX(2, 1, 2, 128, 1) RUNX(2, 1, 3, 128, 1) RUNX(2, 1, 4, 128, 1) RUNX(4, 1, 2, 128, 3) RUNX(4, 1, 3, 128, 3) RUNX(4, 1, 4, 128, 3)
        return 0;
    }
    if (argc > 2 && atoi(argv[2]) == 9) {
        RUNX(2, 1, 0, 128, 1) RUNX(2, 2, 0, 128, 1) RUNX(2, 3, 0, 128, 1) RUNX(2, 4, 0, 128, 1) RUNX(2, 6, 0, 128, 1) RUNX(2, 12, 0, 128, 1)
        RUNX(2, 1, 1, 128, 1) RUNX(2, 1, 2, 128, 1) RUNX(2, 12, 1, 128, 1) RUNX(2, 12, 2, 128, 1)
        RUNX(4, 1, 0, 128, 3) RUNX(4, 2, 0, 128, 3) RUNX(4, 3, 0, 128, 3) RUNX(4, 1, 1, 128, 3) RUNX(4, 1, 2, 128, 3)
        RUNX(2, 2, 0, 128, 6) RUNX(2, 3, 0, 128, 6) RUNX(2, 4, 0, 128, 5) RUNX(2, 2, 0, 64, 10) RUNX(2, 2, 0, 256, 3)
        return 0;
    }
    const int only = argc > 2 ? atoi(argv[2]) : -1;     // profile a single variant under ncu
    if (only == 0) { RUN(2, 2, 128, 1) return 0; }
    if (only == 1) { RUN(4, 2, 128, 3) return 0; }
    if (only == 2) { RUNS(2, 128, 1) return 0; }
    RUN(1, 1, 128, 1) RUN(1, 2, 128, 1)
    RUN(2, 1, 128, 1) RUN(2, 2, 128, 1) RUN(2, 2, 64, 1) RUN(2, 2, 256, 1) RUN(2, 2, 128, 6)
    RUN(4, 1, 128, 1) RUN(4, 2, 128, 1) RUN(4, 2, 64, 1) RUN(4, 2, 128, 3)
    RUNP(1, 128, 1) RUNP(2, 128, 1) RUNP(2, 128, 4) RUNP(2, 128, 5) RUNP(2, 64, 8) RUNP(4, 128, 1) RUNP(4, 128, 2) RUNP(4, 128, 3) RUNP(4, 64, 4) RUNP(4, 64, 6)
    RUNS(2, 128, 1)
    // correctness cross-check between two variants
    { int ng = B * V / 2; k_conv<2, 2, 128, 1><<<(ng + 127) / 128, 128>>>(p, in, out, ng, L); ng = B * V / 4; k_conv_pf<4, 128, 3><<<(ng + 127) / 128, 128>>>(p, in, out2, ng, L); }
    std::vector<double> a(n), b2(n);
    cudaMemcpy(a.data(), out, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b2.data(), out2, n * 8, cudaMemcpyDeviceToHost);
    double md = 0; for (size_t i = 0; i < n; ++i) md = fmax(md, fabs(a[i] - b2[i]));
    printf("max diff between variants %.3e  (%s)\n", md, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
