// Field-transformed HMC on sm_100a (fp64): the 8-subset CNN field transformation, its Jacobian
// log-det and the hand-derived reverse pass (no autograd), behind the C ABI of include/u1hmc.h.
//
// Reference semantics: 2d_u1_cluster/field_trans_opt.py:110-383, cnn_model_opt.py:6-247,
// hmc_u1_ft.py:77-209,348-358, utils.py:21-141.
//
// Structure per subset k (theta^(k) -> theta^(k+1)):
//   k_ft_features : X[6] = sin/cos of the masked plaquette / rectangle angles
//   k_conv<...>   : 3x3 circular convolutions, one launch per layer, weights travel as a
//                   __grid_constant__ launch parameter so every DFMA reads its weight from the
//                   constant bank; epilogue = bias (+skip) + exact-erf GELU (+ GELU' stored for
//                   the reverse pass) or the final atan/2pi
//   k_ft_update   : Delta / log(1+J) on the active links
// Reverse pass (k = 7..0): k_ft_update_bwd -> transposed convolutions (same kernel, flipped
// weights, epilogue (conv + skip) * GELU') -> k_ft_feat_bwd -> k_ft_scatter.
// The FT path is FP64-FMA bound; activations of a chunk of chains are sized to stay in L2.
#include <new>
#include <vector>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include "u1_common.cuh"

namespace u1 {

constexpr int kW = 12;               // hidden width of every jointCNN_* (cnn_model_opt.py)
constexpr int kXC = 6;               // input feature channels
constexpr int kConvThreads = 128;

template <int CIN, int COUT>
struct ConvP {
    double w[CIN * 9 * COUT];        // w[(c*9 + tap)*COUT + o]
    double b[COUT];
};

enum { EPI_GELU = 0, EPI_FINAL = 1, EPI_BWD = 2 };

// ---------------------------------------------------------------------------------------
// exact-erf GELU (nn.GELU default, a = z*Phi(z)) and its derivative (Phi + z*phi) from a piecewise
// Taylor table held in shared memory, evaluated by one Horner pass that carries the derivative.
// The table is built in 80-bit arithmetic at handle creation (build_gelu_table).
//   U1_GELU_V == 1: table of Phi on [-8.5, 8.5], 137 intervals of width 1/8, degree 11
//                   (truncation < 2e-20 for Phi, < 4e-18 for phi): 22 DFMA + 12 LDS per value.
//   U1_GELU_V == 2: table of the odd part Psi = Phi - 1/2 on [0, 8.5] only (Phi(z) = 1/2 + sgn(z) Psi(|z|),
//                   phi even): 137 intervals of width 1/16, degree 9 (truncation < 1e-18 for Psi,
//                   < 3e-17 for phi): 17 DFMA + 10 LDS per value, |z| and the clamp done on the integer
//                   pipe (no fmin/fmax), absolute accuracy like 0.5*(1+erf) itself.
//   U1_GELU_V == 3: the same odd-part table with 273 intervals of width 1/32 and degree 7 (truncation < 1e-19 for
//                   Psi, < 1e-17 for phi; 19.7 KB of shared memory instead of 12 KB): 13 DFMA + 8 LDS per value.
//                   Against 40-digit references on 40 000 points in [-9, 9] all three variants have the same maximum
//                   absolute error, 1.8e-15 (GELU) / 7.3e-16 (GELU'), i.e. the rounding of z*Phi itself.
// |z| > 8.5 is clamped: Phi(-8.5) = 9.5e-18, phi(8.5) = 8e-17.
// ---------------------------------------------------------------------------------------
#ifndef U1_GELU_V
#define U1_GELU_V 3
#endif
#if U1_GELU_V == 1
constexpr int kGeluIntervals = 137, kGeluDeg = 11, kGeluStride = 13;   // odd stride: conflict-free banks
constexpr double kGeluZmax = 8.5, kGeluInvH = 8.0;
#elif U1_GELU_V == 2
constexpr int kGeluIntervals = 137, kGeluDeg = 9, kGeluStride = 11;
constexpr double kGeluZmax = 8.5, kGeluInvH = 16.0;
#else
constexpr int kGeluIntervals = 273, kGeluDeg = 7, kGeluStride = 9;
constexpr double kGeluZmax = 8.5, kGeluInvH = 32.0;
#endif
constexpr int kGeluTableDoubles = (kGeluIntervals * kGeluStride + 1) & ~1;   // even: moved in 16-byte chunks

#if U1_GELU_V == 1
// Interval i is centred at i/8 (i = -68..68, row i+68 of the table); i = rint(8 z) comes for free from
// the low word of 8z + 1.5*2^52.
__device__ __forceinline__ void gelu_and_grad(const double* __restrict__ tab, double z, double& a, double& d) {
    constexpr double kMagic = 6755399441055744.0;
    const double zc = fmin(fmax(z, -kGeluZmax), kGeluZmax);
    const double m = fma(zc, kGeluInvH, kMagic);
    const int i = __double2loint(m);
    const double t = fma(m - kMagic, -1.0 / kGeluInvH, zc);
    const double* q = tab + (i + 68) * kGeluStride;
    double p = q[kGeluDeg], dp = 0.0;
#pragma unroll
    for (int k = kGeluDeg - 1; k >= 0; --k) {
        dp = fma(dp, t, p);
        p = fma(p, t, q[k]);
    }
    a = z * p;
    d = fma(z, dp, p);
}

static void build_gelu_table(double* out) {
    const long double sqrt2pi = sqrtl(2.0L * acosl(-1.0L));
    for (int i = 0; i < kGeluIntervals; ++i) {
        const long double c = ((long double)i - 68.0L) / (long double)kGeluInvH;
        const long double phi = expl(-0.5L * c * c) / sqrt2pi;
        long double he_prev = 0.0L, he = 1.0L;          // He_{-1} := 0, He_0 = 1
        long double fact = 1.0L, sign = 1.0L;
        out[i * kGeluStride + 0] = (double)(0.5L * erfcl(-c / sqrtl(2.0L)));
        for (int k = 1; k <= kGeluDeg; ++k) {            // a_k = (-1)^(k-1) He_{k-1}(c) phi(c) / k!
            fact *= (long double)k;
            out[i * kGeluStride + k] = (double)(sign * he * phi / fact);
            const long double he_next = c * he - (long double)(k - 1) * he_prev;
            he_prev = he; he = he_next; sign = -sign;
        }
        for (int k = kGeluDeg + 1; k < kGeluStride; ++k) out[i * kGeluStride + k] = 0.0;
    }
}
#else
// Interval i is centred at i/InvH (InvH = 16 or 32); i = rint(InvH |z|) is the low word of InvH|z| + 1.5*2^52.
// |z| is clamped to 8.5 by an unsigned min on the high word (doubles >= 0 order like their bit patterns).
__device__ __forceinline__ void gelu_and_grad(const double* __restrict__ tab, double z, double& a, double& d) {
    constexpr double kMagic = 6755399441055744.0;
    constexpr unsigned kHiMax = 0x40210000u;             // high word of 8.5
    const unsigned zh = (unsigned)__double2hiint(z);
    const unsigned ah = zh & 0x7fffffffu;
    const bool big = ah >= kHiMax;
    const double az = __hiloint2double((int)(big ? kHiMax : ah), big ? 0 : __double2loint(z));
    const double m = fma(az, kGeluInvH, kMagic);
    const int i = __double2loint(m);
    const double t = fma(m - kMagic, -1.0 / kGeluInvH, az);
    const double* q = tab + i * kGeluStride;
    double p = q[kGeluDeg];
    double dp = p;
    p = fma(p, t, q[kGeluDeg - 1]);
#pragma unroll
    for (int k = kGeluDeg - 2; k >= 0; --k) {
        dp = fma(dp, t, p);
        p = fma(p, t, q[k]);
    }
    // p = Psi(|z|), dp = phi(|z|);  Phi(z) = 1/2 + sgn(z) Psi(|z|)
    const double ps = __hiloint2double(__double2hiint(p) ^ (int)(zh & 0x80000000u), __double2loint(p));
    const double Phi = 0.5 + ps;
    a = z * Phi;
    d = fma(z, dp, Phi);
}

static void build_gelu_table(double* out) {
    const long double sqrt2pi = sqrtl(2.0L * acosl(-1.0L));
    for (int i = 0; i < kGeluIntervals; ++i) {
        const long double c = (long double)i / (long double)kGeluInvH;
        const long double phi = expl(-0.5L * c * c) / sqrt2pi;
        long double he_prev = 0.0L, he = 1.0L;          // He_{-1} := 0, He_0 = 1
        long double fact = 1.0L, sign = 1.0L;
        out[i * kGeluStride + 0] = (double)(0.5L * erfl(c / sqrtl(2.0L)));          // Psi(c)
        for (int k = 1; k <= kGeluDeg; ++k) {            // a_k = (-1)^(k-1) He_{k-1}(c) phi(c) / k!
            fact *= (long double)k;
            out[i * kGeluStride + k] = (double)(sign * he * phi / fact);
            const long double he_next = c * he - (long double)(k - 1) * he_prev;
            he_prev = he; he = he_next; sign = -sign;
        }
        for (int k = kGeluDeg + 1; k < kGeluStride; ++k) out[i * kGeluStride + k] = 0.0;
    }
}
#endif

// out[o](x,y) = sum_{c,i,j} w[c][i*3+j][o] * in[c]((x+i-1)%L, (y+j-1)%L)   (+ epilogue)
// One thread = two vertically adjacent sites (x, x+1; x even): a 4x3 window per input channel feeds
// 2*9*COUT DFMAs; the channel loop stays rolled (constant-bank weights indexed by a uniform register)
// so the body fits the instruction cache.
//   EPI_GELU : z = acc + b (+ add);  out = GELU(z);  dout = GELU'(z)            (dout may be null)
//   EPI_FINAL: z = acc + b;  y = GELU(z);  out = atan(y)/pi/2;  dout = GELU'(z)/((1+y^2) 2pi)
//   EPI_BWD  : out = (acc (+ add)) (* mul)
template <int CIN, int COUT, int EPI>
__global__ void __launch_bounds__(kConvThreads, 5)      // <=102 registers: forward+GELU 100 us vs 116 us uncapped (microbench)
k_conv(const __grid_constant__ ConvP<CIN, COUT> p, const double* __restrict__ in,
       const double* __restrict__ add, const double* __restrict__ mul,
       double* __restrict__ out, double* __restrict__ dout, const double* __restrict__ gelu_tab,
       int npairs, int L) {
    extern __shared__ double s_tab[];
    // Programmatic dependent launch: let the next kernel of the stream start its prologue while this
    // grid drains; everything before griddepcontrol.wait touches only launch-invariant data.
    asm volatile("griddepcontrol.launch_dependents;");
    if (EPI != EPI_BWD) {
        for (int i = threadIdx.x; i < kGeluTableDoubles; i += kConvThreads) s_tab[i] = gelu_tab[i];
        __syncthreads();
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int s = blockIdx.x * kConvThreads + threadIdx.x;
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
                for (int o = 0; o < COUT; ++o) {
                    const double wv = w[(i * 3 + j) * COUT + o];
                    a0[o] = fma(wv, v[i][j], a0[o]);
                    a1[o] = fma(wv, v[i + 1][j], a1[o]);
                }
    }
    const size_t ob = (size_t)b * COUT * V + (size_t)x * L + y;
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
        const size_t i0 = ob + (size_t)o * V, i1 = i0 + L;
        if (EPI == EPI_BWD) {
            double r0 = a0[o], r1 = a1[o];
            if (add) { r0 += add[i0]; r1 += add[i1]; }
            if (mul) { r0 *= mul[i0]; r1 *= mul[i1]; }
            out[i0] = r0;
            out[i1] = r1;
        } else {
            double z0 = a0[o] + p.b[o], z1 = a1[o] + p.b[o];
            if (EPI == EPI_GELU && add) { z0 += add[i0]; z1 += add[i1]; }
            double g0, d0, g1, d1;
            gelu_and_grad(s_tab, z0, g0, d0);
            gelu_and_grad(s_tab, z1, g1, d1);
            if (EPI == EPI_GELU) {
                out[i0] = g0;
                out[i1] = g1;
                if (dout) { dout[i0] = d0; dout[i1] = d1; }
            } else {
                out[i0] = atan(g0) / kPi / 2;                      // cnn_model_opt.py:46
                out[i1] = atan(g1) / kPi / 2;
                if (dout) {
                    dout[i0] = d0 / ((1.0 + g0 * g0) * kTwoPi);
                    dout[i1] = d1 / ((1.0 + g1 * g1) * kTwoPi);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Shared-memory tiled variant for L in {32, 64, 128}: a CTA of 128 threads owns 128/L groups of ROWS
// complete lattice rows; per input channel the (ROWS*128/L + 2) x L tile (one halo row above and
// below, columns wrap inside the row) is staged by cp.async into a 3-deep ring, so the next two
// channels stream in while the current one feeds the DFMAs and the window reads become LDS.
// Inner loop measured at 72.9 % (ROWS=2) / 76.6 % (ROWS=4) of the DFMA peak against 64 % for the
// direct-load kernel (tools/microbench/conv_variants.cu, profiles/conv_variants5.log).  Both directions
// use ROWS=4 at 128 registers (minBlocks 4).  Forward layers used ROWS=2 with <= 102 registers while the
// GELU epilogue was expensive; with the v2 table and the skip operand at the accumulator start ROWS=4 is
// 1.5 % faster in the whole force (interleaved A/B, profiles/r01_ft_variants_epilogue.log).  ROWS=2 needs
// minBlocks 5 or 6 (<= 102 registers): at 128 registers ptxas schedules that body 30 % slower.
// Same epilogues and launch protocol as k_conv.
// ---------------------------------------------------------------------------------------
constexpr int kTileStages = 3;
#ifndef U1_TILE_FWD_MINB
#define U1_TILE_FWD_MINB 4
#endif
#ifndef U1_TILE_BWD_MINB
#define U1_TILE_BWD_MINB 4
#endif
#ifndef U1_TILE_FWD_ROWS
#define U1_TILE_FWD_ROWS 4
#endif
// Epilogue variants measured on B200 (profiles/r01_ft_variants_epilogue.log): an L2 prefetch (CCTL.E.PF2) of the
// epilogue operands and row-batched operand loads were both slower and are gone; the two switches below
// remain as A/B knobs (defaults = the faster setting).
#ifndef U1_ADD_AT_START      // skip operand folded into the accumulator start values: 1 forward layers, 2 reverse layers too
#define U1_ADD_AT_START 2
#endif
#ifndef U1_TAB_CPASYNC       // GELU table staged by cp.async (joins the first tile's group)
#define U1_TAB_CPASYNC 1
#endif

// ROWS < 4 (small batches, see tile_rows_for) needs the tighter register cap: at 128 registers ptxas schedules that
// body ~30 % slower (round-1 measurement).
template <int LL, int ROWS, int CIN, int COUT, int EPI>
__global__ void __launch_bounds__(kConvThreads, (ROWS < 4 ? 6 : (ROWS > 4 ? 3 : (EPI == EPI_BWD ? U1_TILE_BWD_MINB : U1_TILE_FWD_MINB))))
k_conv_tile(const __grid_constant__ ConvP<CIN, COUT> p, const double* __restrict__ in,
            const double* __restrict__ add, const double* __restrict__ mul,
            double* __restrict__ out, double* __restrict__ dout, const double* __restrict__ gelu_tab) {
    constexpr int TP = kConvThreads / LL;             // row groups per CTA
    constexpr int TROWS = ROWS * TP + 2;              // tile rows incl. halo
    constexpr int TILE = TROWS * LL;                  // doubles per channel tile
    constexpr int V = LL * LL;
    extern __shared__ __align__(16) double s_dyn[];
    double* tile = s_dyn;                             // [kTileStages][TILE]
    double* s_tab = s_dyn + kTileStages * TILE;       // GELU table (forward epilogues only)
    asm volatile("griddepcontrol.launch_dependents;");
    if (EPI != EPI_BWD) {                             // table rides in the first cp.async group (not needed before the epilogue)
#if U1_TAB_CPASYNC
        for (int i = 2 * threadIdx.x; i < kGeluTableDoubles; i += 2 * kConvThreads) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(s_tab + i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gelu_tab + i));
        }
#else
        for (int i = threadIdx.x; i < kGeluTableDoubles; i += kConvThreads) s_tab[i] = gelu_tab[i];
#endif
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    constexpr int groups_per_chain = V / ROWS;
    const int s0 = blockIdx.x * kConvThreads;
    const int b = s0 / groups_per_chain, rp0 = s0 - b * groups_per_chain;
    const int xg0 = rp0 / LL;
    const int tid = threadIdx.x;
    const int tp = tid / LL, y = tid - tp * LL;
    const int x = ROWS * (xg0 + tp);
    const double* ib = in + (size_t)b * CIN * V;
    auto issue = [&](int c, int stage) {
        const double* pc = ib + (size_t)c * V;
        constexpr int CHUNKS = TILE / 2;
        for (int q = tid; q < CHUNKS; q += kConvThreads) {
            const int r = (2 * q) / LL, col = 2 * q - r * LL;
            int xx = ROWS * xg0 - 1 + r;
            xx = xx < 0 ? xx + LL : (xx >= LL ? xx - LL : xx);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + stage * TILE + r * LL + col);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(pc + xx * LL + col));
        }
        asm volatile("cp.async.commit_group;");
    };
#pragma unroll
    for (int st = 0; st < kTileStages - 1; ++st) issue(st, st);
    const int ym = (y == 0 ? LL - 1 : y - 1), yp = (y == LL - 1 ? 0 : y + 1);
    const int r0 = ROWS * tp;
    const size_t ob = (size_t)b * COUT * V + (size_t)x * LL + y;
    // Forward layers start from bias (+ skip connection): the skip operand comes from DRAM, and loading it
    // here puts its latency under the wait for the first tile instead of exposing it once per value in
    // the epilogue (U1_ADD_AT_START=0: old placement).
    double acc[ROWS][COUT];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            acc[r][o] = (EPI == EPI_BWD) ? 0.0 : p.b[o];
            if (U1_ADD_AT_START && (EPI != EPI_BWD || U1_ADD_AT_START > 1) && add) acc[r][o] += add[ob + (size_t)o * V + (size_t)r * LL];
        }
#pragma unroll 1
    for (int c = 0; c < CIN; ++c) {
        if (c + kTileStages - 1 <= CIN) asm volatile("cp.async.wait_group %0;" ::"n"(kTileStages - 2));
        else asm volatile("cp.async.wait_group 0;");
        __syncthreads();                              // tile c visible; buffer of tile c-1 free (also covers s_tab)
        if (c + kTileStages - 1 < CIN) issue(c + kTileStages - 1, (c + kTileStages - 1) % kTileStages);
        const double* t = tile + (c % kTileStages) * TILE;
        double v[ROWS + 2][3];
#pragma unroll
        for (int i = 0; i < ROWS + 2; ++i) {
            v[i][0] = t[(r0 + i) * LL + ym];
            v[i][1] = t[(r0 + i) * LL + y];
            v[i][2] = t[(r0 + i) * LL + yp];
        }
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
    // Epilogue.  The skip operand is already inside acc (U1_ADD_AT_START); GELU' of the reverse layers is
    // loaded here, interleaved by the compiler with the stores of the previous values.
    constexpr bool kAddHere = !(U1_ADD_AT_START && (EPI != EPI_BWD || U1_ADD_AT_START > 1));
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            const size_t idx = ob + (size_t)o * V + (size_t)r * LL;
            if (EPI == EPI_BWD) {
                double rv = acc[r][o];
                if (kAddHere && add) rv += add[idx];
                if (mul) rv *= mul[idx];
                out[idx] = rv;
            } else {
                double z = acc[r][o];
                if (kAddHere && add) z += add[idx];
                double g, d;
                gelu_and_grad(s_tab, z, g, d);
                out[idx] = g;
                if (dout) dout[idx] = d;
            }
        }
}

// ---------------------------------------------------------------------------------------
// Lattice angle helpers (utils.py:118-141) with periodic wrap; x,y may be off by up to 2.
// ---------------------------------------------------------------------------------------
struct Lat {
    const double* t0; const double* t1; int L;
    __device__ __forceinline__ double a0(int x, int y) const { return t0[wrapi(x, L) * L + wrapi(y, L)]; }
    __device__ __forceinline__ double a1(int x, int y) const { return t1[wrapi(x, L) * L + wrapi(y, L)]; }
    __device__ __forceinline__ double P(int x, int y) const {
        return a0(x, y) - a1(x, y) - a0(x, y + 1) + a1(x + 1, y);
    }
    __device__ __forceinline__ double R0(int x, int y) const {
        return a0(x, y) + a0(x + 1, y) + a1(x + 2, y) - a0(x + 1, y + 1) - a0(x, y + 1) - a1(x, y);
    }
    __device__ __forceinline__ double R1(int x, int y) const {
        return a0(x, y) + a1(x + 1, y) + a1(x + 1, y + 1) - a0(x, y + 2) - a1(x, y + 1) - a1(x, y);
    }
    __device__ __forceinline__ int loc(int x, int y) const { return wrapi(x, L) * L + wrapi(y, L); }
};

struct Subset { int mu, a, b; };
__host__ __device__ __forceinline__ Subset subset_of(int k) {   // utils.py:21-50
    Subset s; s.mu = k >> 2; s.a = (k >> 1) & 1; s.b = k & 1; return s;
}

// The six angles of the active link (mu; x,y), their signs in Delta, their dense K channel and
// their location (field_trans_opt.py:137-217).  Angles 0,1 are plaquettes, 2..5 rectangles.
struct LinkTerms { double A[6]; int loc[6]; };
__device__ __forceinline__ void link_terms(const Lat& g, int mu, int x, int y, LinkTerms& t) {
    if (mu == 0) {
        t.A[0] = g.P(x, y);          t.loc[0] = g.loc(x, y);
        t.A[1] = g.P(x, y - 1);      t.loc[1] = g.loc(x, y - 1);
        t.A[2] = g.R0(x - 1, y);     t.loc[2] = g.loc(x - 1, y);
        t.A[3] = g.R0(x - 1, y - 1); t.loc[3] = g.loc(x - 1, y - 1);
        t.A[4] = g.R0(x, y);         t.loc[4] = g.loc(x, y);
        t.A[5] = g.R0(x, y - 1);     t.loc[5] = g.loc(x, y - 1);
    } else {
        t.A[0] = g.P(x, y);          t.loc[0] = g.loc(x, y);
        t.A[1] = g.P(x - 1, y);      t.loc[1] = g.loc(x - 1, y);
        t.A[2] = g.R1(x, y - 1);     t.loc[2] = g.loc(x, y - 1);
        t.A[3] = g.R1(x - 1, y - 1); t.loc[3] = g.loc(x - 1, y - 1);
        t.A[4] = g.R1(x, y);         t.loc[4] = g.loc(x, y);
        t.A[5] = g.R1(x - 1, y);     t.loc[5] = g.loc(x - 1, y);
    }
}
// sign of term j in Delta: mu=0: -,+,-,+,-,+ ; mu=1: +,-,+,-,+,-
__device__ __forceinline__ double term_sign(int mu, int j) { return ((j & 1) ^ mu) ? 1.0 : -1.0; }
// dense K channel of term j: K0 -> 0..3, K1 -> 4..11
__device__ __forceinline__ int term_channel(int mu, int j) {
    return (j < 2) ? (2 * mu + j) : (4 + 4 * mu + (j - 2));
}

// ---------------------------------------------------------------------------------------
// Final convolution restricted to what the subset update reads: K is used only on the active sites
// (1/4 of the lattice) and only in the 6 channels of the active direction (field_trans_opt.py:
// 137-217 multiplies everything else by the field mask), so the last layer costs 1/8 of a dense one
// in both directions.  Results on the entries that are read are identical to the dense layer.
// ---------------------------------------------------------------------------------------
constexpr int kActCh = 6;

// forward: one thread per active site; p holds the 6 needed output channels of the final layer.
__global__ void __launch_bounds__(kConvThreads)
k_conv_final_sparse(const __grid_constant__ ConvP<kW, kActCh> p, const double* __restrict__ in,
                    double* __restrict__ Kout, double* __restrict__ dout,
                    const double* __restrict__ gelu_tab, int k, int nact, int L) {
    extern __shared__ double s_tab[];
    asm volatile("griddepcontrol.launch_dependents;");
    for (int i = threadIdx.x; i < kGeluTableDoubles; i += kConvThreads) s_tab[i] = gelu_tab[i];
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int t = blockIdx.x * kConvThreads + threadIdx.x;
    if (t >= nact) return;
    const int V = L * L, H = L >> 1, QV = V >> 2;
    const Subset ss = subset_of(k);
    const int b = t / QV, q = t - b * QV;
    const int xa = q / H, ya = q - xa * H;
    const int x = 2 * xa + ss.a, y = 2 * ya + ss.b;
    const int xr[3] = {(x == 0 ? L - 1 : x - 1) * L, x * L, (x == L - 1 ? 0 : x + 1) * L};
    const int yc[3] = {(y == 0 ? L - 1 : y - 1), y, (y == L - 1 ? 0 : y + 1)};
    const double* ib = in + (size_t)b * kW * V;
    double acc[kActCh];
#pragma unroll
    for (int j = 0; j < kActCh; ++j) acc[j] = 0.0;
#pragma unroll 1
    for (int c = 0; c < kW; ++c) {
        const double* pc = ib + (size_t)c * V;
        double v[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) v[i * 3 + j] = pc[xr[i] + yc[j]];
        const double* w = p.w + c * 9 * kActCh;
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
#pragma unroll
            for (int j = 0; j < kActCh; ++j) acc[j] = fma(w[tp * kActCh + j], v[tp], acc[j]);
    }
    const size_t ob = (size_t)b * kW * V + (size_t)x * L + y;
#pragma unroll
    for (int j = 0; j < kActCh; ++j) {
        const size_t idx = ob + (size_t)term_channel(ss.mu, j) * V;
        double g, d;
        gelu_and_grad(s_tab, acc[j] + p.b[j], g, d);
        Kout[idx] = atan(g) / kPi / 2;                                   // cnn_model_opt.py:46
        if (dout) dout[idx] = d / ((1.0 + g * g) * kTwoPi);
    }
}

// reverse: hbar = convT_final(zbar) * mul, zbar non-zero only on active sites / active channels.
// One thread per 2x2 block whose top-left corner is an active site; pT = transposed final weights.
__global__ void __launch_bounds__(kConvThreads)
k_conv_final_bwd_sparse(const __grid_constant__ ConvP<kW, kW> pT, const double* __restrict__ zbar,
                        const double* __restrict__ mul, double* __restrict__ out, int k, int nblk, int L) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int t = blockIdx.x * kConvThreads + threadIdx.x;
    if (t >= nblk) return;
    const int V = L * L, H = L >> 1, QV = V >> 2;
    const Subset ss = subset_of(k);
    const int b = t / QV, q = t - b * QV;
    const int xa = q / H, ya = q - xa * H;
    const int X0 = 2 * xa + ss.a, Y0 = 2 * ya + ss.b;
    const int X1 = wrapi(X0 + 1, L), X2 = wrapi(X0 + 2, L), Y1 = wrapi(Y0 + 1, L), Y2 = wrapi(Y0 + 2, L);
    const int src[4] = {X0 * L + Y0, X0 * L + Y2, X2 * L + Y0, X2 * L + Y2};
    const double* zb = zbar + (size_t)b * kW * V;
    double a00[kW], a01[kW], a10[kW], a11[kW];
#pragma unroll
    for (int c = 0; c < kW; ++c) { a00[c] = 0.0; a01[c] = 0.0; a10[c] = 0.0; a11[c] = 0.0; }
#pragma unroll 1
    for (int j = 0; j < kActCh; ++j) {
        const int o = term_channel(ss.mu, j);
        const double* zo = zb + (size_t)o * V;
        const double z00 = zo[src[0]], z01 = zo[src[1]], z10 = zo[src[2]], z11 = zo[src[3]];
        const double* w = pT.w + o * 9 * kW;          // w[tap'*12 + c], tap' = position of the source relative to the site
#pragma unroll
        for (int c = 0; c < kW; ++c) {
            a00[c] = fma(w[4 * kW + c], z00, a00[c]);
            a01[c] = fma(w[3 * kW + c], z00, a01[c]);
            a01[c] = fma(w[5 * kW + c], z01, a01[c]);
            a10[c] = fma(w[1 * kW + c], z00, a10[c]);
            a10[c] = fma(w[7 * kW + c], z10, a10[c]);
            a11[c] = fma(w[0 * kW + c], z00, a11[c]);
            a11[c] = fma(w[2 * kW + c], z01, a11[c]);
            a11[c] = fma(w[6 * kW + c], z10, a11[c]);
            a11[c] = fma(w[8 * kW + c], z11, a11[c]);
        }
    }
    const size_t ob = (size_t)b * kW * V;
    const int s00 = X0 * L + Y0, s01 = X0 * L + Y1, s10 = X1 * L + Y0, s11 = X1 * L + Y1;
#pragma unroll
    for (int c = 0; c < kW; ++c) {
        const size_t cb = ob + (size_t)c * V;
        out[cb + s00] = a00[c] * mul[cb + s00];
        out[cb + s01] = a01[c] * mul[cb + s01];
        out[cb + s10] = a10[c] * mul[cb + s10];
        out[cb + s11] = a11[c] * mul[cb + s11];
    }
}

// ---------------------------------------------------------------------------------------
// Features (field_trans_opt.py:110-131)
// ---------------------------------------------------------------------------------------
__global__ void k_ft_features(const double* __restrict__ th, double* __restrict__ X, int k, int nsites, int L) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    const Subset ss = subset_of(k);
    Lat g{th + (size_t)b * 2 * V, th + (size_t)b * 2 * V + V, L};
    const bool active_site = ((x & 1) == ss.a) && ((y & 1) == ss.b);
    const bool mP = ss.mu == 0 ? ((x & 1) != ss.a) : ((y & 1) != ss.b);      // utils.py:52-80
    const bool mR = !active_site;                                             // utils.py:82-115
    double sP = 0.0, cP = 1.0, sR = 0.0, cR = 1.0;
    if (mP) fast_sincos(g.P(x, y), sP, cP);
    if (mR) fast_sincos(ss.mu == 0 ? g.R1(x, y) : g.R0(x, y), sR, cR);
    double* o = X + (size_t)b * kXC * V + r;
    o[0] = sP;
    o[(size_t)V] = cP;
    o[(size_t)2 * V] = ss.mu == 0 ? 0.0 : sR;     // sin(m0 R0)
    o[(size_t)3 * V] = ss.mu == 0 ? sR : 0.0;     // sin(m1 R1)
    o[(size_t)4 * V] = ss.mu == 0 ? 1.0 : cR;     // cos(m0 R0)
    o[(size_t)5 * V] = ss.mu == 0 ? cR : 1.0;     // cos(m1 R1)
}

// ---------------------------------------------------------------------------------------
// Subset update: theta^(k+1) = theta^(k) + Delta_k, log-det partial sums
// (field_trans_opt.py:137-217, 284-383).  mode 0: out = theta + Delta ; mode 1: out = Delta.
// ---------------------------------------------------------------------------------------
constexpr int kUpdThreads = 256;

__global__ void __launch_bounds__(kUpdThreads)
k_ft_update(const double* __restrict__ th, const double* __restrict__ K, double* __restrict__ out,
            double* __restrict__ partial, int k, int L, int nblk, int mode) {
    __shared__ double red[32];
    const int V = L * L;
    const int b = blockIdx.y;
    const int r = blockIdx.x * kUpdThreads + threadIdx.x;
    const Subset ss = subset_of(k);
    double ld[1] = {0.0};
    if (r < V) {
        const int x = r / L, y = r - x * L;
        const double* t0 = th + (size_t)b * 2 * V;
        Lat g{t0, t0 + V, L};
        double o0 = mode ? 0.0 : t0[r], o1 = mode ? 0.0 : t0[V + r];
        if (((x & 1) == ss.a) && ((y & 1) == ss.b)) {
            LinkTerms t;
            link_terms(g, ss.mu, x, y, t);
            const double* Kb = K + (size_t)b * kW * V + r;
            double d[6], j[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                double sn, cs;
                fast_sincos(t.A[q], sn, cs);
                const double kq = Kb[(size_t)term_channel(ss.mu, q) * V];
                d[q] = kq * (term_sign(ss.mu, q) * sn);
                j[q] = kq * (-cs);
            }
            // same association as the reference: (plaq0+plaq1) + (((r0+r1)+r2)+r3)
            const double delta = (d[0] + d[1]) + (((d[2] + d[3]) + d[4]) + d[5]);
            const double jac = (j[0] + j[1]) + (((j[2] + j[3]) + j[4]) + j[5]);
            if (ss.mu == 0) o0 += delta; else o1 += delta;
            ld[0] = log(1.0 + jac);
        }
        double* ob = out + (size_t)b * 2 * V;
        ob[r] = o0;
        ob[V + r] = o1;
    }
    if (partial) {
        block_sum<1>(ld, red);
        if (threadIdx.x == 0) partial[(size_t)b * nblk + blockIdx.x] = ld[0];
    }
}

// Final convolution + subset update in one kernel (forward sweep of the sampler): the thread that evaluates the six K
// channels of an active site is the only consumer of those values in k_ft_update, so it goes on to form Delta and
// log(1+J) itself and writes theta^(k+1) for its 2x2 block of sites (the active site and its three inactive
// neighbours), instead of K making a round trip through memory and a second kernel re-reading theta.  Kout / dout
// (needed by the reverse pass) are written only when dout != nullptr.  Same arithmetic, in the same order, as
// k_conv_final_sparse followed by k_ft_update (mode 0); only the grouping of the log-det partial sums differs
// (128 active sites per CTA instead of 256 sites).  Needs (L*L/4) % 128 == 0, i.e. L a multiple of 32.
__global__ void __launch_bounds__(kConvThreads)
k_final_update(const __grid_constant__ ConvP<kW, kActCh> p, const double* __restrict__ in,
               const double* __restrict__ th, double* __restrict__ th_out,
               double* __restrict__ Kout, double* __restrict__ dout, double* __restrict__ partial,
               const double* __restrict__ gelu_tab, int k, int L) {
    extern __shared__ double s_tab[];
    __shared__ double red[32];
    asm volatile("griddepcontrol.launch_dependents;");
    for (int i = threadIdx.x; i < kGeluTableDoubles; i += kConvThreads) s_tab[i] = gelu_tab[i];
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int t = blockIdx.x * kConvThreads + threadIdx.x;
    const int V = L * L, H = L >> 1, QV = V >> 2;
    const Subset ss = subset_of(k);
    const int b = t / QV, q = t - b * QV;
    const int xa = q / H, ya = q - xa * H;
    const int x = 2 * xa + ss.a, y = 2 * ya + ss.b;
    const int xr[3] = {(x == 0 ? L - 1 : x - 1) * L, x * L, (x == L - 1 ? 0 : x + 1) * L};
    const int yc[3] = {(y == 0 ? L - 1 : y - 1), y, (y == L - 1 ? 0 : y + 1)};
    const double* ib = in + (size_t)b * kW * V;
    double acc[kActCh];
#pragma unroll
    for (int j = 0; j < kActCh; ++j) acc[j] = 0.0;
#pragma unroll 1
    for (int c = 0; c < kW; ++c) {
        const double* pc = ib + (size_t)c * V;
        double v[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) v[i * 3 + j] = pc[xr[i] + yc[j]];
        const double* w = p.w + c * 9 * kActCh;
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
#pragma unroll
            for (int j = 0; j < kActCh; ++j) acc[j] = fma(w[tp * kActCh + j], v[tp], acc[j]);
    }
    const double* t0 = th + (size_t)b * 2 * V;
    Lat g{t0, t0 + V, L};
    LinkTerms lt;
    link_terms(g, ss.mu, x, y, lt);
    const int r = x * L + y;
    const size_t ob = (size_t)b * kW * V + r;
    double d[6], jj[6];
#pragma unroll
    for (int j = 0; j < kActCh; ++j) {
        double gg, dd;
        gelu_and_grad(s_tab, acc[j] + p.b[j], gg, dd);
        const double kq = atan(gg) / kPi / 2;                               // cnn_model_opt.py:46
        if (dout) {
            const size_t idx = ob + (size_t)term_channel(ss.mu, j) * V;
            Kout[idx] = kq;
            dout[idx] = dd / ((1.0 + gg * gg) * kTwoPi);
        }
        double sn, cs;
        fast_sincos(lt.A[j], sn, cs);
        d[j] = kq * (term_sign(ss.mu, j) * sn);
        jj[j] = kq * (-cs);
    }
    // same association as the reference: (plaq0+plaq1) + (((r0+r1)+r2)+r3)
    const double delta = (d[0] + d[1]) + (((d[2] + d[3]) + d[4]) + d[5]);
    const double jac = (jj[0] + jj[1]) + (((jj[2] + jj[3]) + jj[4]) + jj[5]);
    double ld[1] = {log(1.0 + jac)};
    // theta^(k+1) of the 2x2 block {x, x^1} x {y, y^1} (x^1: the other site of the pair 2xa, 2xa+1)
    double* o0 = th_out + (size_t)b * 2 * V;
    const int xo = (x ^ 1) * L, yo = y ^ 1;
    const int site[4] = {r, x * L + yo, xo + y, xo + yo};
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        double a0 = t0[site[m]], a1 = t0[V + site[m]];
        if (m == 0) { if (ss.mu == 0) a0 += delta; else a1 += delta; }
        o0[site[m]] = a0;
        o0[V + site[m]] = a1;
    }
    if (partial) {
        block_sum<1>(ld, red);
        const int nblk = QV / kConvThreads;
        if (threadIdx.x == 0) partial[(size_t)b * nblk + (blockIdx.x - b * nblk)] = ld[0];
    }
}

// k_final_update with the input staged like k_conv_tile (L in {32, 64, 128}): the CTA owns 4*128/LL lattice rows, i.e.
// 128 active sites, one per thread; per input channel the rows (plus one halo row above and below) come in by cp.async
// through the 3-deep ring, so the nine window reads per channel are LDS instead of dependent global loads (ncu of the
// direct-load form: long_scoreboard 6.0 of 11 stall cycles per issue, 75 us at 342 chains for 30 us of HBM traffic).
// Same arithmetic in the same order as k_final_update.
template <int LL>
__global__ void __launch_bounds__(kConvThreads)
k_final_update_tile(const __grid_constant__ ConvP<kW, kActCh> p, const double* __restrict__ in,
                    const double* __restrict__ th, double* __restrict__ th_out,
                    double* __restrict__ Kout, double* __restrict__ dout, double* __restrict__ partial,
                    const double* __restrict__ gelu_tab, int k) {
    constexpr int TP = kConvThreads / LL, NR = 4 * TP, TROWS = NR + 2, TILE = TROWS * LL, V = LL * LL, HC = LL / 2;
    constexpr int NBLK = LL / NR;                       // CTAs per chain
    extern __shared__ __align__(16) double s_dyn[];
    __shared__ double red[32];
    double* tile = s_dyn;                               // [kTileStages][TILE]
    double* s_tab = s_dyn + kTileStages * TILE;
    asm volatile("griddepcontrol.launch_dependents;");
    for (int i = 2 * threadIdx.x; i < kGeluTableDoubles; i += 2 * kConvThreads) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_tab + i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gelu_tab + i));
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x;
    const int b = blockIdx.x / NBLK, R0 = (blockIdx.x - b * NBLK) * NR;
    const Subset ss = subset_of(k);
    const int ar = tid / HC, ac = tid - ar * HC;
    const int x = R0 + 2 * ar + ss.a, y = 2 * ac + ss.b;
    const double* ib = in + (size_t)b * kW * V;
    auto issue = [&](int c, int stage) {
        const double* pc = ib + (size_t)c * V;
        constexpr int CHUNKS = TILE / 2;
        for (int q = tid; q < CHUNKS; q += kConvThreads) {
            const int r = (2 * q) / LL, col = 2 * q - r * LL;
            int xx = R0 - 1 + r;
            xx = xx < 0 ? xx + LL : (xx >= LL ? xx - LL : xx);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + stage * TILE + r * LL + col);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(pc + xx * LL + col));
        }
        asm volatile("cp.async.commit_group;");
    };
#pragma unroll
    for (int st = 0; st < kTileStages - 1; ++st) issue(st, st);
    const int ym = (y == 0 ? LL - 1 : y - 1), yp = (y == LL - 1 ? 0 : y + 1);
    const int tr = x - R0;                              // tile rows tr, tr+1, tr+2 = lattice rows x-1, x, x+1
    // theta of the 2x2 block and the six angles while the first tiles are in flight
    const double* t0 = th + (size_t)b * 2 * V;
    Lat g{t0, t0 + V, LL};
    LinkTerms lt;
    link_terms(g, ss.mu, x, y, lt);
    const int r = x * LL + y;
    const int xo = (x ^ 1) * LL, yo = y ^ 1;
    const int site[4] = {r, x * LL + yo, xo + y, xo + yo};
    double o0v[4], o1v[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) { o0v[m] = t0[site[m]]; o1v[m] = t0[V + site[m]]; }
    double acc[kActCh];
#pragma unroll
    for (int j = 0; j < kActCh; ++j) acc[j] = 0.0;
#pragma unroll 1
    for (int c = 0; c < kW; ++c) {
        if (c + kTileStages - 1 <= kW) asm volatile("cp.async.wait_group %0;" ::"n"(kTileStages - 2));
        else asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        if (c + kTileStages - 1 < kW) issue(c + kTileStages - 1, (c + kTileStages - 1) % kTileStages);
        const double* t = tile + (c % kTileStages) * TILE;
        double v[9];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            v[i * 3 + 0] = t[(tr + i) * LL + ym];
            v[i * 3 + 1] = t[(tr + i) * LL + y];
            v[i * 3 + 2] = t[(tr + i) * LL + yp];
        }
        const double* w = p.w + c * 9 * kActCh;
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
#pragma unroll
            for (int j = 0; j < kActCh; ++j) acc[j] = fma(w[tp * kActCh + j], v[tp], acc[j]);
    }
    const size_t ob = (size_t)b * kW * V + r;
    double d[6], jj[6];
#pragma unroll
    for (int j = 0; j < kActCh; ++j) {
        double gg, dd;
        gelu_and_grad(s_tab, acc[j] + p.b[j], gg, dd);
        const double kq = atan(gg) / kPi / 2;                               // cnn_model_opt.py:46
        if (dout) {
            const size_t idx = ob + (size_t)term_channel(ss.mu, j) * V;
            Kout[idx] = kq;
            dout[idx] = dd / ((1.0 + gg * gg) * kTwoPi);
        }
        double sn, cs;
        fast_sincos(lt.A[j], sn, cs);
        d[j] = kq * (term_sign(ss.mu, j) * sn);
        jj[j] = kq * (-cs);
    }
    const double delta = (d[0] + d[1]) + (((d[2] + d[3]) + d[4]) + d[5]);
    const double jac = (jj[0] + jj[1]) + (((jj[2] + jj[3]) + jj[4]) + jj[5]);
    double ld[1] = {log(1.0 + jac)};
    double* o0 = th_out + (size_t)b * 2 * V;
    if (ss.mu == 0) o0v[0] += delta; else o1v[0] += delta;
#pragma unroll
    for (int m = 0; m < 4; ++m) { o0[site[m]] = o0v[m]; o0[V + site[m]] = o1v[m]; }
    if (partial) {
        block_sum<1>(ld, red);
        if (threadIdx.x == 0) partial[(size_t)b * NBLK + (blockIdx.x - b * NBLK)] = ld[0];
    }
}

// One fixed-point iteration of the inverse transformation (field_trans_opt.py:245-282):
// next = curr - ft_phase(iter); per-block partial sums of |next-iter|^2 and |iter|^2.
__global__ void __launch_bounds__(kUpdThreads)
k_ft_inverse_iter(const double* __restrict__ cur, const double* __restrict__ it, const double* __restrict__ K,
                  double* __restrict__ nxt, double* __restrict__ partial, int k, int L, int nblk) {
    __shared__ double red[2 * 32];
    const int V = L * L;
    const int b = blockIdx.y;
    const int r = blockIdx.x * kUpdThreads + threadIdx.x;
    const Subset ss = subset_of(k);
    double v[2] = {0.0, 0.0};
    if (r < V) {
        const int x = r / L, y = r - x * L;
        const double* i0 = it + (size_t)b * 2 * V;
        const double* c0 = cur + (size_t)b * 2 * V;
        double n0 = c0[r], n1 = c0[V + r];
        if (((x & 1) == ss.a) && ((y & 1) == ss.b)) {
            Lat g{i0, i0 + V, L};
            LinkTerms t;
            link_terms(g, ss.mu, x, y, t);
            const double* Kb = K + (size_t)b * kW * V + r;
            double d[6];
#pragma unroll
            for (int q = 0; q < 6; ++q)
                d[q] = Kb[(size_t)term_channel(ss.mu, q) * V] * (term_sign(ss.mu, q) * sin(t.A[q]));
            const double delta = (d[0] + d[1]) + (((d[2] + d[3]) + d[4]) + d[5]);
            if (ss.mu == 0) n0 = n0 + (-delta); else n1 = n1 + (-delta);
        }
        const double a0 = i0[r], a1 = i0[V + r];
        v[0] = (n0 - a0) * (n0 - a0) + (n1 - a1) * (n1 - a1);
        v[1] = a0 * a0 + a1 * a1;
        double* ob = nxt + (size_t)b * 2 * V;
        ob[r] = n0;
        ob[V + r] = n1;
    }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) {
        partial[2 * ((size_t)b * nblk + blockIdx.x)] = v[0];
        partial[2 * ((size_t)b * nblk + blockIdx.x) + 1] = v[1];
    }
}

// norms[0..1] = sums of n partial pairs, single CTA, fixed order.
__global__ void __launch_bounds__(256) k_sum_pairs(const double* __restrict__ partial, size_t n, double* __restrict__ norms) {
    __shared__ double red[2 * 32];
    double v[2] = {0.0, 0.0};
    for (size_t i = threadIdx.x; i < n; i += 256) { v[0] += partial[2 * i]; v[1] += partial[2 * i + 1]; }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) { norms[0] = v[0]; norms[1] = v[1]; }
}

// logdet[b] = sum over the 8 subsets (in order) of the fixed-order sum of that subset's block partials.
// partial: [8][B][nblk].  Same association as accumulating subset by subset.
__global__ void k_ft_logdet_accum(const double* __restrict__ partial, double* __restrict__ logdet, int B, int nblk) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) {
        const double* p = partial + ((size_t)k * B + b) * nblk;
        double s = 0.0;
        for (int i = 0; i < nblk; ++i) s += p[i];
        tot = (k == 0) ? s : tot + s;
    }
    logdet[b] = tot;
}

// ---------------------------------------------------------------------------------------
// Reverse pass of one subset.
//   g      : dS/dtheta^(k+1)                         [B][2][V]
//   zbar   : adjoint of the final conv's pre-activation  [B][12][V]  (written densely)
//   AP, AR : adjoints of the plaquette / (update-type) rectangle angles, [B][V] (AR: every entry written;
//            AP: written on the complement of the CNN plaquette mask, the rest is never read)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kUpdThreads)
k_ft_update_bwd(const double* __restrict__ th, const double* __restrict__ K,
                const double* __restrict__ dfin, const double* __restrict__ g,
                double* __restrict__ zbar, double* __restrict__ AP, double* __restrict__ AR,
                int k, int nsites, int L) {
    const int s = blockIdx.x * kUpdThreads + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    const Subset ss = subset_of(k);
    if (((x & 1) != ss.a) || ((y & 1) != ss.b)) return;      // zbar is consumed on active sites only
    {
        const double* t0 = th + (size_t)b * 2 * V;
        Lat gl{t0, t0 + V, L};
        LinkTerms t;
        link_terms(gl, ss.mu, x, y, t);
        const double* Kb = K + (size_t)b * kW * V + r;
        const double* db = dfin + (size_t)b * kW * V + r;
        double* zo = zbar + (size_t)b * kW * V + r;
        double sn[6], cs[6], kq[6], jq[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            fast_sincos(t.A[q], sn[q], cs[q]);
            kq[q] = Kb[(size_t)term_channel(ss.mu, q) * V];
            jq[q] = kq[q] * (-cs[q]);
        }
        const double jac = (jq[0] + jq[1]) + (((jq[2] + jq[3]) + jq[4]) + jq[5]);
        const double ga = g[(size_t)b * 2 * V + (size_t)ss.mu * V + r];   // adjoint of Delta
        const double gb = -1.0 / (1.0 + jac);                             // adjoint of J  (S_FT has -log(1+J))
        double* APb = AP + (size_t)b * V;
        double* ARb = AR + (size_t)b * V;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double sg = term_sign(ss.mu, q);
            const int ch = term_channel(ss.mu, q);
            const double kbar = ga * sg * sn[q] - gb * cs[q];
            zo[(size_t)ch * V] = kbar * db[(size_t)ch * V];
            const double abar = kq[q] * (ga * sg * cs[q] + gb * sn[q]);
            if (q < 2) APb[t.loc[q]] = abar; else ARb[t.loc[q]] = abar;
        }
    }
}

// k_ft_update_bwd + k_conv_final_bwd_sparse in one kernel (reverse sweep of the sampler).  A CTA owns TAX = 128/H
// rows of active sites (H = L/2 active sites per row, all of them: the y direction wraps inside the CTA) and has one
// extra warp group for the row below: (TAX+1)*H threads each form zbar (6 channels) of ONE active site exactly as
// k_ft_update_bwd does and park it in shared memory; the TAX*H owner threads also write AP / AR.  After the barrier
// the owner threads run the transposed final convolution of their 2x2 output block from the four zbar entries in
// shared memory - same operations in the same order as k_conv_final_bwd_sparse, so the result is bit-identical -
// and zbar itself never goes to memory.  H in {16, 32, 64}.  (Two passes of six output channels with the GELU' factors
// requested ahead of the FMAs were measured slower: 165 vs 112 us at 342 chains, profiles/r02_ft_fusion_ab.log.)
template <int H>
__global__ void __launch_bounds__(kConvThreads + H)
k_final_bwd_fused(const __grid_constant__ ConvP<kW, kW> pT, const double* __restrict__ th, const double* __restrict__ K,
                  const double* __restrict__ dfin, const double* __restrict__ g, const double* __restrict__ mul,
                  double* __restrict__ out, double* __restrict__ AP, double* __restrict__ AR, int k) {
    constexpr int L = 2 * H, V = L * L, TAX = kConvThreads / H, NBLK = H / TAX;      // NBLK CTAs per chain
    __shared__ double zs[kActCh][(TAX + 1) * H];
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const Subset ss = subset_of(k);
    const int b = blockIdx.x / NBLK, xa0 = (blockIdx.x - b * NBLK) * TAX;
    const int tid = threadIdx.x, rx = tid / H, ya = tid - rx * H;
    const bool owner = rx < TAX;
    const int xa = (xa0 + rx) % H;
    const int X0 = 2 * xa + ss.a, Y0 = 2 * ya + ss.b;
    {
        const int r = X0 * L + Y0;
        const double* t0 = th + (size_t)b * 2 * V;
        Lat gl{t0, t0 + V, L};
        LinkTerms t;
        link_terms(gl, ss.mu, X0, Y0, t);
        const double* Kb = K + (size_t)b * kW * V + r;
        const double* db = dfin + (size_t)b * kW * V + r;
        double sn[6], cs[6], kq[6], jq[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            fast_sincos(t.A[q], sn[q], cs[q]);
            kq[q] = Kb[(size_t)term_channel(ss.mu, q) * V];
            jq[q] = kq[q] * (-cs[q]);
        }
        const double jac = (jq[0] + jq[1]) + (((jq[2] + jq[3]) + jq[4]) + jq[5]);
        const double ga = g[(size_t)b * 2 * V + (size_t)ss.mu * V + r];   // adjoint of Delta
        const double gb = -1.0 / (1.0 + jac);                             // adjoint of J  (S_FT has -log(1+J))
        double* APb = AP + (size_t)b * V;
        double* ARb = AR + (size_t)b * V;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double sg = term_sign(ss.mu, q);
            const int ch = term_channel(ss.mu, q);
            const double kbar = ga * sg * sn[q] - gb * cs[q];
            zs[q][tid] = kbar * db[(size_t)ch * V];
            if (owner) {
                const double abar = kq[q] * (ga * sg * cs[q] + gb * sn[q]);
                if (q < 2) APb[t.loc[q]] = abar; else ARb[t.loc[q]] = abar;
            }
        }
    }
    __syncthreads();
    if (!owner) return;
    const int yb = (ya + 1 == H) ? 0 : ya + 1;
    const int X1 = wrapi(X0 + 1, L), Y1 = wrapi(Y0 + 1, L);
    // the 24 zbar values of this 2x2 block in registers, then one rolled loop over the 12 output channels: 54 FMAs per
    // channel against constant-bank weights indexed by the channel counter, the four GELU' factors of the NEXT channel
    // requested before the FMAs of the current one (few registers, so many CTAs per SM hide the store latency).
    // Per accumulator the j = 0..5 order is that of k_conv_final_bwd_sparse: bit-identical.
    double z00[kActCh], z01[kActCh], z10[kActCh], z11[kActCh];
#pragma unroll
    for (int j = 0; j < kActCh; ++j) {
        z00[j] = zs[j][rx * H + ya];       z01[j] = zs[j][rx * H + yb];
        z10[j] = zs[j][(rx + 1) * H + ya]; z11[j] = zs[j][(rx + 1) * H + yb];
    }
    const size_t ob = (size_t)b * kW * V;
    const int s00 = X0 * L + Y0, s01 = X0 * L + Y1, s10 = X1 * L + Y0, s11 = X1 * L + Y1;
    int oc[kActCh];
#pragma unroll
    for (int j = 0; j < kActCh; ++j) oc[j] = term_channel(ss.mu, j) * 9 * kW;
    double m00 = mul[ob + s00], m01 = mul[ob + s01], m10 = mul[ob + s10], m11 = mul[ob + s11];
#pragma unroll 1
    for (int c = 0; c < kW; ++c) {
        const size_t cb = ob + (size_t)c * V;
        double n00 = 0.0, n01 = 0.0, n10 = 0.0, n11 = 0.0;
        if (c + 1 < kW) { n00 = mul[cb + V + s00]; n01 = mul[cb + V + s01]; n10 = mul[cb + V + s10]; n11 = mul[cb + V + s11]; }
        double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
#pragma unroll
        for (int j = 0; j < kActCh; ++j) {
            const double* w = pT.w + oc[j] + c;           // w[tap'*12 + c], tap' = position of the source relative to the site
            a00 = fma(w[4 * kW], z00[j], a00);
            a01 = fma(w[3 * kW], z00[j], a01);
            a01 = fma(w[5 * kW], z01[j], a01);
            a10 = fma(w[1 * kW], z00[j], a10);
            a10 = fma(w[7 * kW], z10[j], a10);
            a11 = fma(w[0 * kW], z00[j], a11);
            a11 = fma(w[2 * kW], z01[j], a11);
            a11 = fma(w[6 * kW], z10[j], a11);
            a11 = fma(w[8 * kW], z11[j], a11);
        }
        out[cb + s00] = a00 * m00;
        out[cb + s01] = a01 * m01;
        out[cb + s10] = a10 * m10;
        out[cb + s11] = a11 * m11;
        m00 = n00; m01 = n01; m10 = n10; m11 = n11;
    }
}

// Total plaquette adjoint PT[b][V] of subset k: update angles (AP, AR) + CNN features (Xbar).
__global__ void k_ft_feat_bwd(const double* __restrict__ th, const double* __restrict__ Xbar,
                              const double* __restrict__ AP, const double* __restrict__ AR,
                              double* __restrict__ PT, int k, int nsites, int L) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    const Subset ss = subset_of(k);
    const double* t0 = th + (size_t)b * 2 * V;
    Lat g{t0, t0 + V, L};
    const double* xb = Xbar + (size_t)b * kXC * V;
    const double* ARb = AR + (size_t)b * V;
    // AP (adjoint of the update's plaquette angles) is written by k_ft_update_bwd exactly on the plaquettes that
    // touch an active link - the complement of the CNN plaquette mask mP - and AR on every site, so neither
    // buffer needs zeroing between subsets: unwritten AP entries are simply not read.
    const bool mP = ss.mu == 0 ? ((x & 1) != ss.a) : ((y & 1) != ss.b);
    double pt = mP ? 0.0 : AP[(size_t)b * V + r];
    if (mP) {
        double sn, cs;
        fast_sincos(g.P(x, y), sn, cs);
        pt += xb[r] * cs - xb[(size_t)V + r] * sn;
    }
    // rectangle adjoint of the CNN-feature type at location (xx,yy)
    auto cnn_rect = [&](int xx, int yy) -> double {
        const int xw = wrapi(xx, L), yw = wrapi(yy, L);
        if (((xw & 1) == ss.a) && ((yw & 1) == ss.b)) return 0.0;       // masked out
        const int q = xw * L + yw;
        double sn, cs;
        if (ss.mu == 0) {            // features use R1: channels 3 (sin), 5 (cos)
            fast_sincos(g.R1(xw, yw), sn, cs);
            return xb[(size_t)3 * V + q] * cs - xb[(size_t)5 * V + q] * sn;
        } else {                     // features use R0: channels 2 (sin), 4 (cos)
            fast_sincos(g.R0(xw, yw), sn, cs);
            return xb[(size_t)2 * V + q] * cs - xb[(size_t)4 * V + q] * sn;
        }
    };
    // R0(x,y)=P(x,y)+P(x+1,y) -> P(x,y) collects R0bar(x,y)+R0bar(x-1,y);
    // R1(x,y)=P(x,y)+P(x,y+1) -> P(x,y) collects R1bar(x,y)+R1bar(x,y-1).
    if (ss.mu == 0) {   // update angles are R0 (AR), CNN angles are R1
        pt += ARb[r] + ARb[g.loc(x - 1, y)];
        pt += cnn_rect(x, y) + cnn_rect(x, y - 1);
    } else {            // update angles are R1 (AR), CNN angles are R0
        pt += ARb[r] + ARb[g.loc(x, y - 1)];
        pt += cnn_rect(x, y) + cnn_rect(x - 1, y);
    }
    PT[(size_t)b * V + r] = pt;
}

// g_out = g_in + dP/dtheta^T PT:  th0(x,y): +PT(x,y) - PT(x,y-1);  th1(x,y): -PT(x,y) + PT(x-1,y).
__global__ void k_ft_scatter(const double* __restrict__ gin, const double* __restrict__ PT,
                             double* __restrict__ gout, int nsites, int L) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    const double* p = PT + (size_t)b * V;
    const double c = p[r];
    const double l = p[x * L + wrapi(y - 1, L)];
    const double u = p[wrapi(x - 1, L) * L + y];
    const size_t o = (size_t)b * 2 * V + r;
    gout[o] = gin[o] + (c - l);
    gout[o + V] = gin[o + V] + (u - c);
}

// y = y + a*x (elementwise)
__global__ void k_axpy(double* __restrict__ y, const double* __restrict__ x, double a, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = fma(a, x[i], y[i]);
}
__global__ void k_axpy_wrap(double* __restrict__ y, const double* __restrict__ x, double a, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = wrap_pi(fma(a, x[i], y[i]));
}

// sums3[b] = {sum cos(wrap P), sum wrap P, sum pi^2}; one CTA per chain (L <= ~256 in practice).
__global__ void __launch_bounds__(256)
k_chain_sums(const double* __restrict__ th, const double* __restrict__ pi, double* __restrict__ sums, int L) {
    __shared__ double red[3 * 32];
    const int V = L * L, b = blockIdx.x;
    const double* t0 = th + (size_t)b * 2 * V;
    Lat g{t0, t0 + V, L};
    double v[3] = {0.0, 0.0, 0.0};
    for (int r = threadIdx.x; r < V; r += blockDim.x) {
        const int x = r / L, y = r - x * L;
        const double w = wrap_pi(g.P(x, y));
        v[0] += cos(w);
        v[1] += w;
        if (pi) {
            const double a = pi[(size_t)b * 2 * V + r], c = pi[(size_t)b * 2 * V + V + r];
            v[2] += a * a + c * c;
        }
    }
    block_sum<3>(v, red);
    if (threadIdx.x == 0) { sums[b * 3] = v[0]; sums[b * 3 + 1] = v[1]; sums[b * 3 + 2] = v[2]; }
}

// plain force with the reference's wrap (initial adjoint dS/dtheta^(8)), cf. u1_plain.cu k_force.
__global__ void k_plain_force(const double* __restrict__ th, double* __restrict__ F, int nsites, int L, double beta) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    const int V = L * L;
    const int b = s / V, r = s - b * V, x = r / L, y = r - x * L;
    const double* t0 = th + (size_t)b * 2 * V;
    Lat g{t0, t0 + V, L};
    const double sp = fast_sin(g.P(x, y));                    // sin(wrap P) = sin P
    F[(size_t)b * 2 * V + r] = beta * (sp - fast_sin(g.P(x, y - 1)));
    F[(size_t)b * 2 * V + V + r] = beta * (-sp + fast_sin(g.P(x - 1, y)));
}

// H = -beta*sums[0] - logdet + 0.5*sums[2]  -> Hout[b]; also S_ft
__global__ void k_ft_hamiltonian(const double* __restrict__ sums, const double* __restrict__ logdet,
                                 double* __restrict__ S_ft, double* __restrict__ H, int B, double beta) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double s = (-beta) * sums[b * 3] - logdet[b];
    if (S_ft) S_ft[b] = s;
    if (H) H[b] = s + 0.5 * sums[b * 3 + 2];
}

__global__ void k_ft_metropolis(const double* __restrict__ H_old, const double* __restrict__ H_new,
                                const double* __restrict__ sums_old, const double* __restrict__ sums_new,
                                const double* __restrict__ u_inject, uint64_t seed, uint64_t chain0,
                                uint64_t traj, int B, double inv_vol, double* out_dH, int32_t* out_acc,
                                double* out_H, double* out_plaq, double* out_topo, int32_t* acc_flag) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double dH = H_new[b] - H_old[b];
    const double u = u_inject ? u_inject[b] : philox_accept_uniform(chain_key(seed, chain0 + b), traj);
    const bool acc = u < exp(-dH);
    acc_flag[b] = acc ? 1 : 0;
    const double* f = (acc ? sums_new : sums_old) + b * 3;
    if (out_dH) out_dH[b] = dH;
    if (out_acc) out_acc[b] = acc ? 1 : 0;
    if (out_H) out_H[b] = acc ? H_new[b] : H_old[b];
    if (out_plaq) out_plaq[b] = f[0] * inv_vol;
    if (out_topo) out_topo[b] = floor(0.1 + f[1] / kTwoPi);
}

// dst[b] = src[b] where acc[b] (per-chain select), n elements per chain
__global__ void k_select_chain(double* __restrict__ dst, const double* __restrict__ src,
                               const int32_t* __restrict__ acc, size_t per_chain, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (acc[i / per_chain]) dst[i] = src[i];
}

__global__ void k_momenta_ft(double* __restrict__ pi, uint64_t seed, uint64_t chain0, uint64_t traj, int nsites, int V) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsites) return;
    const int b = s / V, site = s - b * V;
    double n0, n1;
    philox_site_normals(chain_key(seed, chain0 + b), traj, (uint32_t)site, kLogTab, n0, n1);
    pi[(size_t)b * 2 * V + site] = n0;
    pi[(size_t)b * 2 * V + V + site] = n1;
}

}  // namespace u1

using namespace u1;

// ---------------------------------------------------------------------------------------
// Host side: handle, workspace carving, orchestration
// ---------------------------------------------------------------------------------------
constexpr int kLanes = 4;      // internal streams of u1ft_trajectory (at most)

struct u1ft_handle_s {
    uint32_t magic;
    int n_res, n_layers;                                  // n_layers = 2 + 2*n_res
    std::vector<ConvP<kXC, kW>> f0;                       // [8] first layer (6->12)
    std::vector<ConvP<kW, kW>> f;                         // [8][n_layers-1]
    std::vector<ConvP<kW, kXC>> b0;                       // [8] transposed first layer (12->6)
    std::vector<ConvP<kW, kW>> bt;                        // [8][n_layers-1] transposed
    std::vector<ConvP<kW, kActCh>> fs;                    // [8] final layer, 6 active output channels of subset k
    double* gelu_tab;                                     // device copy of the GELU Taylor table
    int device;
    cudaStream_t lane[kLanes];                            // internal streams of u1ft_trajectory (lazy)
    cudaEvent_t ev_fork, ev_join[kLanes];
    bool lanes_ready;
    // CUDA-graph cache of the trajectory building blocks (one FT force + kick: ~170 kernel nodes with their
    // programmatic-dependent-launch edges; one forward + action), keyed by workspace / shape / parameters
    struct GraphEntry {
        int kind; void* ws; int C, L; double beta, dt;
        cudaGraphExec_t exec; long long kernels; unsigned long long stamp;
    };
    std::vector<GraphEntry> graphs;
    std::mutex graph_mu;
    unsigned long long graph_clock;
    bool graphs_broken;                                   // a capture failed once: stay on eager launches
    cudaStream_t cap_stream;                              // side stream used only for capturing (lazy)
};

static int ensure_lanes(u1ft_handle_s* h) {
    if (h->lanes_ready) return U1_OK;
    cudaError_t e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    for (int l = 0; l < kLanes && e == cudaSuccess; ++l) {
        e = cudaStreamCreateWithFlags(&h->lane[l], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join[l], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return (int)e;
    h->lanes_ready = true;
    return U1_OK;
}
constexpr uint32_t kMagic = 0x55314654u;  // "U1FT"

static inline bool handle_ok(u1ft_handle_t h) { return h && h->magic == kMagic; }
// The handle's GELU table lives on the device that was current at u1ft_create: every entry point must
// run on that device (otherwise the kernels would dereference another GPU's memory).
static inline bool handle_on_current_device(u1ft_handle_t h) {
    int d = -1;
    return cudaGetDevice(&d) == cudaSuccess && d == h->device;
}
static inline size_t al(size_t v) { return (v + 255) / 256 * 256; }
static inline int blocks_for(size_t n, int t) { return (int)((n + t - 1) / t); }
static inline int ew_grid_ft(size_t n) {
    size_t g = (n + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

// Default chunk: as many sites as fit a quarter of the device memory at the rnet3 workspace cost
// (~7.4 KB per site: theta^(0..8), features, activations, K and GELU' of 8 subsets x 8 layers), i.e. all 1024
// chains of C4 in one chunk (30 GB of the 180 GB) on a B200.  Measured FT-force throughput grows
// monotonically with the chunk because kernel-boundary ramps and tails amortise (rnet3, L=64: 15.5 TFLOP/s at
// 55 chains, 22.7 at 342, 23.1 at 512, 23.4 at 1024; an attempt to cut chunks at whole-wave sizes instead,
// 370+370+284, was slower than three equal chunks); the convolutions are FP64-bound, so the activations
// streaming through HBM (~290 B per site-layer) are hidden behind the math.
static long long default_chunk_sites() {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || total_b == 0) return 1818624;
    const long long sites = (long long)(total_b / 4 / 7400);
    return sites < 227328 ? 227328 : sites;
}
static std::atomic<long long> g_chunk_sites{-1};
static long long chunk_sites() {
    long long t = g_chunk_sites.load();
    if (t < 0) {
        const char* e = getenv("U1FT_CHUNK_SITES");
        t = e ? atoll(e) : default_chunk_sites();
        if (t < 1) t = 1;
        g_chunk_sites.store(t);
    }
    return t;
}
// Chunks are independent, so the trajectory entry point can run them on two internal streams ("lanes"); U1FT_LANES=1|2
// forces the count, otherwise it is chosen from the batch size (see below).
static int lanes_for(int B, int L) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("U1FT_LANES");
        forced = e ? atoi(e) : -1;
        if (forced > kLanes) forced = kLanes;
    }
    if (B < 2) return 1;
    if (forced >= 1) return forced;
    // Automatic: two lanes when a layer of the whole batch is only a few waves of CTAs, so that one lane's kernels fill
    // the SMs the other lane's kernel leaves idle while it drains (L=64 rnet3, trajectories/s on one B200, 1 -> 2 lanes:
    // 128 chains 351 -> 362, 256 chains 358 -> 374, 512 and 1024 chains unchanged; L=32 LOSES 13-20 %: its kernels are
    // too short).  The launches come from replayed CUDA graphs, so the host launch rate that made two lanes slower in
    // round 1 (18.7 -> 16.4 TFLOP/s at 128 chains) no longer matters.
    return (L >= 64 && (long long)B * L * L <= 256ll * 4096) ? 2 : 1;
}
static std::atomic<bool> g_chunk_explicit{false};
static int chunk_chains(int B, int L) {
    const long long target = chunk_sites();
    long long c = target / ((long long)L * L);
    if (c < 1) c = 1;
    const int nl = lanes_for(B, L);
    long long nchunks = (B + c - 1) / c;
    if (nchunks < nl) nchunks = nl;
    if (nchunks % nl) nchunks += nl - nchunks % nl;
    long long cc = (B + nchunks - 1) / nchunks;
    if (cc > kMaxGridY) cc = kMaxGridY;                 // chains ride in gridDim.y of k_ft_update (small L, huge batches)
    return (int)cc;
}

struct FtWs {
    double* th[9];        // theta^(0..8)                     [C][2][V]
    double* X;            // features / Xbar                  [C][6][V]
    double* A;            // activation ping                  [C][12][V]
    double* Bf;           // activation pong                  [C][12][V]
    double* Kk[8];        // K of each subset                 [C][12][V]
    double* D;            // GELU' per subset per layer       [8][n_layers][C][12][V]
    double* AP; double* AR; double* PT;                      // [C][V]
    double* g0; double* g1;                                   // adjoint ping-pong [C][2][V]
    double* partial;      // logdet partials                  [8][C][nblk]
    double* logdet; double* sums_old; double* sums_new; double* H_old; double* H_new;   // [C](x3)
    double* th_new; double* pi; double* force;                // trajectory state  [C][2][V]
    int32_t* acc;
    size_t bytes;
};

static FtWs carve_ft(void* ws, int n_layers, int C, int L, bool backward) {
    FtWs w{};
    const size_t V = (size_t)L * L;
    const size_t f2 = al((size_t)C * 2 * V * 8), f6 = al((size_t)C * kXC * V * 8), f12 = al((size_t)C * kW * V * 8);
    const size_t f1 = al((size_t)C * V * 8);
    const int nblk = blocks_for(V, kUpdThreads);
    char* p = (char*)ws;
    auto take = [&](size_t n) { double* r = (double*)p; p += n; return r; };
    for (int i = 0; i < 9; ++i) w.th[i] = take(f2);
    w.X = take(f6); w.A = take(f12); w.Bf = take(f12);
    w.partial = take(al((size_t)8 * C * nblk * 8));                   // [8 subsets][C][nblk]
    w.logdet = take(al((size_t)C * 8));
    w.sums_old = take(al((size_t)C * 3 * 8)); w.sums_new = take(al((size_t)C * 3 * 8));
    w.H_old = take(al((size_t)C * 8)); w.H_new = take(al((size_t)C * 8));
    w.acc = (int32_t*)take(al((size_t)C * 4));
    w.th_new = take(f2); w.pi = take(f2); w.force = take(f2);
    if (backward) {
        for (int i = 0; i < 8; ++i) w.Kk[i] = take(f12);
        w.D = take((size_t)8 * n_layers * f12);
        w.AP = take(f1); w.AR = take(f1); w.PT = take(f1);
        w.g0 = take(f2); w.g1 = take(f2);
    } else {
        for (int i = 0; i < 8; ++i) w.Kk[i] = w.A;   // unused
        w.Kk[0] = take(f12);                          // single K buffer reused by every subset
        for (int i = 1; i < 8; ++i) w.Kk[i] = w.Kk[0];
    }
    w.bytes = (size_t)(p - (char*)ws);
    return w;
}

static inline double* Dbuf(const FtWs& w, int n_layers, int C, int L, int k, int layer) {
    const size_t f12 = al((size_t)C * kW * (size_t)L * L * 8);
    return (double*)((char*)w.D + ((size_t)k * n_layers + layer) * f12);
}

static bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("U1FT_PDL"); v = e ? atoi(e) : 1; }
    return v != 0;
}

static void pdl_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int blocks, size_t smem, cudaStream_t st) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}

template <int LL, int ROWS, int CIN, int COUT, int EPI>
static int launch_conv_tile(u1ft_handle_t h, const ConvP<CIN, COUT>& p, const double* in, const double* add,
                            const double* mul, double* out, double* dout, int nsites, cudaStream_t st) {
    constexpr int TILE = (ROWS * (kConvThreads / LL) + 2) * LL;
    constexpr size_t smem = ((size_t)kTileStages * TILE + (EPI == EPI_BWD ? 0 : kGeluTableDoubles)) * sizeof(double);
    static std::atomic<bool> configured[kMaxDevices];      // function attributes are per device
    const int dev_i = current_device();
    if (!configured[dev_i].load(std::memory_order_acquire)) {
        const char* cv = getenv("U1FT_CARVEOUT");
        if (cv) cudaFuncSetAttribute(k_conv_tile<LL, ROWS, CIN, COUT, EPI>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv));
        configured[dev_i].store(true, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    pdl_config(cfg, attr, nsites / (ROWS * kConvThreads), smem, st);
    const double* tab = h->gelu_tab;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_conv_tile<LL, ROWS, CIN, COUT, EPI>, p, in, add, mul, out, dout, tab);
    count_launch();
    return e == cudaSuccess ? U1_OK : (int)e;
}

static int tile_path_enabled() {          // U1FT_TILE: 0 off, 1 reverse layers only, 2 all layers (default)
    static int v = -1;
    if (v < 0) { const char* e = getenv("U1FT_TILE"); v = e ? atoi(e) : 2; }
    return v;
}

// Lattice rows per thread of the tiled convolution.  4 rows (four DFMAs per constant-bank weight read) is the
// fastest body, but a layer then has only nsites/512 CTAs of 4 warps: at 128 chains of L=32 that is 256 CTAs
// for 148 SMs x 4 resident CTAs, i.e. under two warps per scheduler and nothing to hide the LDS / barrier
// latency of the channel loop.  Small launches therefore trade per-thread reuse for parallelism
// (2 rows per thread).  Measured on one B200 at the per-GPU batch of the 8-GPU run (profiles/r02_ft_small_batch_rows.log,
// rnet3, 128 chains): L=32 (256 CTAs at 4 rows): 710 -> 767 trajectories/s with 2 rows, 397 with 1 row; L=64 (1024
// CTAs): 348 / 338 / 119; L=32 with 256 chains (512 CTAs): 1156 -> 1234.  Hence 2 rows below 600 CTAs, never 1.
// U1FT_TILE_ROWS=1|2|4 forces a value.
static int tile_rows_for(int nsites) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("U1FT_TILE_ROWS"); forced = e ? atoi(e) : 0; }
    if (forced == 1 || forced == 2 || forced == 4) return forced;
    static int t4 = -1, t2 = -1;                      // thresholds in CTAs of the ROWS=4 launch (tunable)
    if (t4 < 0) { const char* e = getenv("U1FT_TILE_T4"); t4 = e ? atoi(e) : 600; }
    if (t2 < 0) { const char* e = getenv("U1FT_TILE_T2"); t2 = e ? atoi(e) : 0; }
    const int ctas4 = nsites / (4 * kConvThreads);
    return ctas4 >= t4 ? 4 : (ctas4 >= t2 ? 2 : 1);
}

static bool rows8_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("U1FT_ROWS8"); v = e ? atoi(e) : 1; }
    return v != 0;
}

template <int CIN, int COUT, int EPI>
static int launch_conv(u1ft_handle_t h, const ConvP<CIN, COUT>& p, const double* in, const double* add,
                       const double* mul, double* out, double* dout, int nsites, int L, cudaStream_t st) {
    // Measured in the full force (342 chains): reverse layers 135 vs 161 us (ncu), forward layers 161 vs 228 us
    // against the direct kernel; the register caps of the tiled kernels are the U1_TILE_*_MINB knobs above.
    if (tile_path_enabled() && (EPI == EPI_BWD || tile_path_enabled() > 1) && (L == 32 || L == 64 || L == 128)) {
        const int rows = tile_rows_for(nsites);
        // 12 -> 6 reverse layer (the first layer transposed): half the FMAs per channel of a 12 -> 12 layer, so 8 rows per
        // thread (48 accumulators again; each constant-bank weight read feeds 8 DFMAs) once the launch is large enough
        if constexpr (EPI == EPI_BWD && COUT == kXC) {
            if (rows == 4 && rows8_enabled() && nsites / (8 * kConvThreads) >= 600) {
                if (L == 64) return launch_conv_tile<64, 8, CIN, COUT, EPI>(h, p, in, add, mul, out, dout, nsites, st);
                if (L == 32) return launch_conv_tile<32, 8, CIN, COUT, EPI>(h, p, in, add, mul, out, dout, nsites, st);
                if (L == 128) return launch_conv_tile<128, 8, CIN, COUT, EPI>(h, p, in, add, mul, out, dout, nsites, st);
            }
        }
#define U1_TILE_CASE(LLv, Rv) \
        if (L == LLv && rows == Rv) return launch_conv_tile<LLv, Rv, CIN, COUT, EPI>(h, p, in, add, mul, out, dout, nsites, st);
        U1_TILE_CASE(64, 4) U1_TILE_CASE(64, 2) U1_TILE_CASE(64, 1)
        U1_TILE_CASE(32, 4) U1_TILE_CASE(32, 2) U1_TILE_CASE(32, 1)
        U1_TILE_CASE(128, 4) U1_TILE_CASE(128, 2) U1_TILE_CASE(128, 1)
#undef U1_TILE_CASE
    }
    const int npairs = nsites >> 1;
    const size_t smem = (EPI == EPI_BWD) ? 0 : (size_t)kGeluTableDoubles * sizeof(double);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    pdl_config(cfg, attr, blocks_for(npairs, kConvThreads), smem, st);
    const double* tab = h->gelu_tab;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_conv<CIN, COUT, EPI>, p, in, add, mul, out, dout, tab, npairs, L);
    count_launch();
    if (e != cudaSuccess) return (int)e;
    return U1_OK;
}

static int launch_final_sparse(u1ft_handle_t h, int k, const double* in, double* Kout, double* dout,
                               int nsites, int L, cudaStream_t st) {
    const int nact = nsites >> 2;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    pdl_config(cfg, attr, blocks_for(nact, kConvThreads), (size_t)kGeluTableDoubles * sizeof(double), st);
    const double* tab = h->gelu_tab;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_conv_final_sparse, h->fs[k], in, Kout, dout, tab, k, nact, L);
    count_launch();
    return e == cudaSuccess ? U1_OK : (int)e;
}

static int launch_final_bwd_sparse(const ConvP<kW, kW>& pT, int k, const double* zbar, const double* mul,
                                   double* out, int nsites, int L, cudaStream_t st) {
    const int nblk = nsites >> 2;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    pdl_config(cfg, attr, blocks_for(nblk, kConvThreads), 0, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_conv_final_bwd_sparse, pT, zbar, mul, out, k, nblk, L);
    count_launch();
    return e == cudaSuccess ? U1_OK : (int)e;
}

// CNN forward of subset k on chunk of C chains: X -> K (w.Kk[k]).  store!=0 keeps GELU' for backward.
static int cnn_forward(u1ft_handle_t h, const FtWs& w, int k, int C, int L, bool store, cudaStream_t st,
                       bool skip_final = false) {
    const int ns = C * L * L, nl = h->n_layers;
    auto D = [&](int layer) { return store ? Dbuf(w, nl, C, L, k, layer) : nullptr; };
    const ConvP<kW, kW>* fk = &h->f[(size_t)k * (nl - 1)];
    U1_TRY((launch_conv<kXC, kW, EPI_GELU>(h, h->f0[k], w.X, nullptr, nullptr, w.A, D(0), ns, L, st)));
    for (int i = 0; i < h->n_res; ++i) {            // ResBlock (cnn_model_opt.py:74-87)
        U1_TRY((launch_conv<kW, kW, EPI_GELU>(h, fk[2 * i], w.A, nullptr, nullptr, w.Bf, D(1 + 2 * i), ns, L, st)));
        U1_TRY((launch_conv<kW, kW, EPI_GELU>(h, fk[2 * i + 1], w.Bf, w.A, nullptr, w.A, D(2 + 2 * i), ns, L, st)));
    }
    (void)fk;
    if (!skip_final) U1_TRY(launch_final_sparse(h, k, w.A, w.Kk[k], D(nl - 1), ns, L, st));     // active sites / channels only
    return U1_OK;
}

// CNN reverse of subset k: zbar (in w.A) -> Xbar (w.X).
static int cnn_backward(u1ft_handle_t h, const FtWs& w, int k, int C, int L, cudaStream_t st, bool skip_final = false) {
    const int ns = C * L * L, nl = h->n_layers;
    auto D = [&](int layer) { return Dbuf(w, nl, C, L, k, layer); };
    const ConvP<kW, kW>* bk = &h->bt[(size_t)k * (nl - 1)];
    // hbar = convT_final(zbar_f); times GELU' of the layer below
    if (!skip_final) U1_TRY(launch_final_bwd_sparse(bk[nl - 2], k, w.A, D(nl - 2), w.Bf, ns, L, st));
    // now w.Bf = vbar of the last block (or zbar0 for `simple`)
    double* cur = w.Bf;  double* oth = w.A;
    for (int i = h->n_res - 1; i >= 0; --i) {
        // z1bar = convT_2(vbar) * GELU'(conv1 pre-activation)
        U1_TRY((launch_conv<kW, kW, EPI_BWD>(h, bk[2 * i + 1], cur, nullptr, D(1 + 2 * i), oth, nullptr, ns, L, st)));
        // vbar_prev = (convT_1(z1bar) + vbar) * GELU'(layer below), in place over vbar
        U1_TRY((launch_conv<kW, kW, EPI_BWD>(h, bk[2 * i], oth, cur, D(2 * i), cur, nullptr, ns, L, st)));
    }
    U1_TRY((launch_conv<kW, kXC, EPI_BWD>(h, h->b0[k], cur, nullptr, nullptr, w.X, nullptr, ns, L, st)));
    return U1_OK;
}

// Reverse-sweep counterpart (k_final_bwd_fused): update adjoint + transposed final convolution, L in {32, 64, 128}.
static bool fuse_final_bwd(int L) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("U1FT_FUSE_BWD"); v = e ? atoi(e) : 1; }
    return v != 0 && (L == 32 || L == 64 || L == 128);
}
template <int H>
static int launch_final_bwd_fused_h(const ConvP<kW, kW>& pT, const double* th, const double* K, const double* dfin,
                                    const double* g, const double* mul, double* out, double* AP, double* AR, int k, int C,
                                    cudaStream_t st) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    pdl_config(cfg, attr, C * (H / (kConvThreads / H)), 0, st);
    cfg.blockDim = dim3(kConvThreads + H);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_final_bwd_fused<H>, pT, th, K, dfin, g, mul, out, AP, AR, k);
    count_launch();
    return e == cudaSuccess ? U1_OK : (int)e;
}
static int launch_final_bwd_fused(const ConvP<kW, kW>& pT, const double* th, const double* K, const double* dfin,
                                  const double* g, const double* mul, double* out, double* AP, double* AR, int k, int C, int L,
                                  cudaStream_t st) {
    if (L == 32) return launch_final_bwd_fused_h<16>(pT, th, K, dfin, g, mul, out, AP, AR, k, C, st);
    if (L == 64) return launch_final_bwd_fused_h<32>(pT, th, K, dfin, g, mul, out, AP, AR, k, C, st);
    if (L == 128) return launch_final_bwd_fused_h<64>(pT, th, K, dfin, g, mul, out, AP, AR, k, C, st);
    return U1_ERR_UNSUPPORTED;
}

// U1FT_FUSE (default 1): final convolution + subset update in one kernel where the lattice allows it.
static bool fuse_final_update(int L) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("U1FT_FUSE"); v = e ? atoi(e) : 1; }
    return v != 0 && ((L * L / 4) % kConvThreads == 0);
}

// Forward sweep over the 8 subsets for a chunk: theta (w.th[0]) -> w.th[8], logdet.
static int ft_forward_chunk(u1ft_handle_t h, const FtWs& w, int C, int L, bool store, bool want_logdet,
                            cudaStream_t st) {
    const int V = L * L, ns = C * V;
    const bool fused = fuse_final_update(L);
    const int nblk = fused ? (V / 4) / kConvThreads : blocks_for(V, kUpdThreads);    // log-det partials per chain and subset
    for (int k = 0; k < 8; ++k) {
        // (computing the features inside the first convolution - six smem tiles formed from theta - was measured
        // and is 0.3 % SLOWER than this separate 20 us kernel: profiles/r02_ft_fusion_ab.log)
        k_ft_features<<<blocks_for(ns, 256), 256, 0, st>>>(w.th[k], w.X, k, ns, L);
        count_launch();
        U1_CHECK_LAUNCH();
        U1_TRY(cnn_forward(h, w, k, C, L, store, st, fused));
        double* part = want_logdet ? w.partial + (size_t)k * C * nblk : nullptr;
        if (fused) {
            cudaLaunchConfig_t cfg;
            cudaLaunchAttribute attr[1];
            pdl_config(cfg, attr, (ns >> 2) / kConvThreads, (size_t)kGeluTableDoubles * sizeof(double), st);
            const double* tab = h->gelu_tab;
            const double* in = w.A;
            const double* th = w.th[k];
            double* dfin = store ? Dbuf(w, h->n_layers, C, L, k, h->n_layers - 1) : nullptr;
            cudaError_t e;
            if (tile_path_enabled() > 1 && (L == 32 || L == 64 || L == 128)) {
                const int tp = kConvThreads / L;
                cfg.dynamicSmemBytes = ((size_t)kTileStages * (4 * tp + 2) * L + kGeluTableDoubles) * sizeof(double);
                if (L == 32) e = cudaLaunchKernelEx(&cfg, k_final_update_tile<32>, h->fs[k], in, th, w.th[k + 1], w.Kk[k], dfin, part, tab, k);
                else if (L == 64) e = cudaLaunchKernelEx(&cfg, k_final_update_tile<64>, h->fs[k], in, th, w.th[k + 1], w.Kk[k], dfin, part, tab, k);
                else e = cudaLaunchKernelEx(&cfg, k_final_update_tile<128>, h->fs[k], in, th, w.th[k + 1], w.Kk[k], dfin, part, tab, k);
            } else {
                e = cudaLaunchKernelEx(&cfg, k_final_update, h->fs[k], in, th, w.th[k + 1], w.Kk[k], dfin, part, tab, k, L);
            }
            count_launch();
            if (e != cudaSuccess) return (int)e;
        } else {
            k_ft_update<<<dim3(nblk, C), kUpdThreads, 0, st>>>(w.th[k], w.Kk[k], w.th[k + 1], part, k, L, nblk, 0);
            count_launch();
            U1_CHECK_LAUNCH();
        }
    }
    if (want_logdet) {
        k_ft_logdet_accum<<<blocks_for(C, 128), 128, 0, st>>>(w.partial, w.logdet, C, nblk);
        count_launch();
        U1_CHECK_LAUNCH();
    }
    return U1_OK;
}

// S_ft (and sums of theta_ori) for the chunk already forwarded.
static int ft_action_chunk(const FtWs& w, int C, int L, double beta, const double* pi, double* sums,
                           double* S_ft, double* H, cudaStream_t st) {
    k_chain_sums<<<C, 256, 0, st>>>(w.th[8], pi, sums, L);
    count_launch();
    U1_CHECK_LAUNCH();
    k_ft_hamiltonian<<<blocks_for(C, 128), 128, 0, st>>>(sums, w.logdet, S_ft, H, C, beta);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

// Reverse sweep: result (dS_FT/dtheta) lands in *result (one of w.g0/w.g1).
static int ft_backward_chunk(u1ft_handle_t h, const FtWs& w, int C, int L, double beta, double** result,
                             cudaStream_t st) {
    const int V = L * L, ns = C * V;
    double* g = w.g0;  double* gn = w.g1;
    k_plain_force<<<blocks_for(ns, 256), 256, 0, st>>>(w.th[8], g, ns, L, beta);
    count_launch();
    U1_CHECK_LAUNCH();
    const bool fused = fuse_final_bwd(L);
    for (int k = 7; k >= 0; --k) {
        const double* dfin = Dbuf(w, h->n_layers, C, L, k, h->n_layers - 1);
        if (fused) {
            const ConvP<kW, kW>& pT = h->bt[(size_t)k * (h->n_layers - 1) + h->n_layers - 2];
            const double* mul = Dbuf(w, h->n_layers, C, L, k, h->n_layers - 2);
            U1_TRY(launch_final_bwd_fused(pT, w.th[k], w.Kk[k], dfin, g, mul, w.Bf, w.AP, w.AR, k, C, L, st));
        } else {
            k_ft_update_bwd<<<blocks_for(ns, kUpdThreads), kUpdThreads, 0, st>>>(w.th[k], w.Kk[k], dfin, g, w.A, w.AP, w.AR, k, ns, L);
            count_launch();
            U1_CHECK_LAUNCH();
        }
        U1_TRY(cnn_backward(h, w, k, C, L, st, fused));
        k_ft_feat_bwd<<<blocks_for(ns, 256), 256, 0, st>>>(w.th[k], w.X, w.AP, w.AR, w.PT, k, ns, L);
        count_launch();
        U1_CHECK_LAUNCH();
        k_ft_scatter<<<blocks_for(ns, 256), 256, 0, st>>>(g, w.PT, gn, ns, L);
        count_launch();
        U1_CHECK_LAUNCH();
        double* t = g; g = gn; gn = t;
    }
    *result = g;
    return U1_OK;
}

static int copy_async(void* dst, const void* src, size_t n, cudaStream_t st) {
    cudaError_t e = cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, st);
    return e == cudaSuccess ? U1_OK : (int)e;
}

static int check_ws_ft(void* ws, size_t have, size_t need) {
    if (!ws || have < need || (reinterpret_cast<uintptr_t>(ws) & 255u)) return U1_ERR_WORKSPACE;
    return U1_OK;
}

// force for a chunk whose theta is in w.th[0]; result pointer returned; optional S_ft output.
static int ft_force_chunk(u1ft_handle_t h, const FtWs& w, int C, int L, double beta, double* S_ft,
                          double** result, cudaStream_t st) {
    U1_TRY(ft_forward_chunk(h, w, C, L, true, true, st));
    if (S_ft) U1_TRY(ft_action_chunk(w, C, L, beta, nullptr, w.sums_new, S_ft, nullptr, st));
    return ft_backward_chunk(h, w, C, L, beta, result, st);
}

// ---------------------------------------------------------------------------------------
// CUDA graphs of the trajectory building blocks.  A 50-step FT trajectory is ~9 000 kernel launches of 10 KB
// parameters each; at small per-GPU batches (128 chains at 8 GPUs, or L=32) the kernels are shorter than the
// ~6 us the host needs per launch, so the GPU starves.  The force + kick of one leapfrog step (and the
// forward + action pass of H_old / H_new) are captured once per (workspace, chunk shape, beta, dt) with their
// programmatic-dependent-launch edges and replayed: ~110 host launches per trajectory instead of ~9 000.
// U1FT_GRAPH=0 keeps eager launches; a stream that is already being captured by the caller is left alone.
// ---------------------------------------------------------------------------------------
thread_local long long g_graph_kernels = 0;      // kernels of this call that ran inside graph replays
thread_local long long g_graph_replays = 0;      // graph launches of this call
enum { GRAPH_FORCE_KICK = 1, GRAPH_ACTION = 2 };
constexpr size_t kMaxGraphs = 12;

static bool graphs_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("U1FT_GRAPH"); v = e ? atoi(e) : 1; }
    return v != 0;
}

// Runs body(stream) - a sequence of kernel launches - either eagerly on st or as a cached graph replayed on st.
// The capture happens on a handle-owned side stream: the caller's stream may be the legacy default stream
// (torch's default), which cannot be captured.
template <class Body>
static int run_graphed(u1ft_handle_s* h, int kind, void* ws, int C, int L, double beta, double dt, cudaStream_t st, Body body) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    const bool usable = graphs_enabled() && !h->graphs_broken && cudaStreamIsCapturing(st, &cs) == cudaSuccess &&
                        cs == cudaStreamCaptureStatusNone;
    if (!usable) { cudaGetLastError(); return body(st); }
    cudaGraphExec_t exec = nullptr;
    long long kernels = 0;
    {
        std::lock_guard<std::mutex> lk(h->graph_mu);
        for (auto& g : h->graphs)
            if (g.kind == kind && g.ws == ws && g.C == C && g.L == L && g.beta == beta && g.dt == dt) {
                g.stamp = ++h->graph_clock;
                exec = g.exec; kernels = g.kernels;
                break;
            }
        if (!exec) {
            auto broken = [&]() { cudaGetLastError(); h->graphs_broken = true; };
            if (!h->cap_stream && cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess) broken();
            const long long k0 = g_launches;
            if (!h->graphs_broken && cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) broken();
            if (!h->graphs_broken) {
                const int r = body(h->cap_stream);             // recorded, not executed
                cudaGraph_t graph = nullptr;
                const cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
                kernels = g_launches - k0;
                g_launches = k0;                                 // the replay below accounts for the kernels
                if (r != U1_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
                    exec = nullptr;
                    broken();
                }
                if (graph) cudaGraphDestroy(graph);
            }
            if (exec) {
                if (h->graphs.size() >= kMaxGraphs) {            // evict the least recently used entry
                    size_t lru = 0;
                    for (size_t i = 1; i < h->graphs.size(); ++i) if (h->graphs[i].stamp < h->graphs[lru].stamp) lru = i;
                    cudaGraphExecDestroy(h->graphs[lru].exec);
                    h->graphs.erase(h->graphs.begin() + lru);
                }
                h->graphs.push_back({kind, ws, C, L, beta, dt, exec, kernels, ++h->graph_clock});
            }
        }
    }
    if (!exec) return body(st);                                  // capture unavailable: eager launches, for good
    const cudaError_t e = cudaGraphLaunch(exec, st);
    if (e != cudaSuccess) return (int)e;
    count_launch((int)kernels);
    g_graph_kernels += kernels;
    g_graph_replays += 1;
    return U1_OK;
}

extern "C" {

long long u1ft_last_host_launch_count(void) { return g_launches - g_graph_kernels + g_graph_replays; }

long long u1ft_set_chunk_sites(long long sites) {
    const long long prev = g_chunk_explicit.load() ? chunk_sites() : 0;     // 0 = automatic (wave-aware) choice
    if (sites < 1) {
        g_chunk_sites.store(-1);
        g_chunk_explicit.store(false);
    } else {
        g_chunk_sites.store(sites);
        g_chunk_explicit.store(true);
    }
    return prev;
}

size_t u1ft_weight_count(int n_res) {
    if (n_res < 0 || n_res > 8) return 0;
    const size_t first = (size_t)kW * kXC * 9 + kW, rest = (size_t)kW * kW * 9 + kW;
    return 8 * (first + (size_t)(1 + 2 * n_res) * rest);
}

int u1ft_create(u1ft_handle_t* out, int n_res, const double* weights, size_t n_doubles) {
    if (!out || !weights) return U1_ERR_ARG;
    if (n_res < 0 || n_res > 8 || n_doubles != u1ft_weight_count(n_res)) return U1_ERR_HANDLE;
    u1ft_handle_s* h = new (std::nothrow) u1ft_handle_s();
    if (!h) return U1_ERR_HANDLE;
    h->magic = kMagic; h->n_res = n_res; h->n_layers = 2 + 2 * n_res;
    h->lanes_ready = false;
    h->graph_clock = 0;
    h->graphs_broken = false;
    h->cap_stream = nullptr;
    const int nl = h->n_layers;
    h->fs.resize(8);
    h->f0.resize(8); h->b0.resize(8); h->f.resize((size_t)8 * (nl - 1)); h->bt.resize((size_t)8 * (nl - 1));
    const double* p = weights;
    for (int s = 0; s < 8; ++s) {
        for (int l = 0; l < nl; ++l) {
            const int cin = (l == 0) ? kXC : kW;
            const double* W = p;  p += (size_t)kW * cin * 9;      // W[o][c][tap]
            const double* bias = p;  p += kW;
            if (l == 0) {
                for (int o = 0; o < kW; ++o) {
                    h->f0[s].b[o] = bias[o];
                    for (int c = 0; c < kXC; ++c)
                        for (int t = 0; t < 9; ++t) {
                            const double v = W[((size_t)o * kXC + c) * 9 + t];
                            h->f0[s].w[(c * 9 + t) * kW + o] = v;
                            h->b0[s].w[(o * 9 + (8 - t)) * kXC + c] = v;     // transposed + flipped
                        }
                }
                for (int c = 0; c < kXC; ++c) h->b0[s].b[c] = 0.0;
            } else {
                ConvP<kW, kW>& F = h->f[(size_t)s * (nl - 1) + (l - 1)];
                ConvP<kW, kW>& T = h->bt[(size_t)s * (nl - 1) + (l - 1)];
                for (int o = 0; o < kW; ++o) {
                    F.b[o] = bias[o];
                    T.b[o] = 0.0;
                    for (int c = 0; c < kW; ++c)
                        for (int t = 0; t < 9; ++t) {
                            const double v = W[((size_t)o * kW + c) * 9 + t];
                            F.w[(c * 9 + t) * kW + o] = v;
                            T.w[(o * 9 + (8 - t)) * kW + c] = v;
                        }
                }
                if (l == nl - 1) {                     // final layer: the 6 channels subset s reads
                    const int mu = s >> 2;
                    for (int j = 0; j < kActCh; ++j) {
                        const int o = (j < 2) ? (2 * mu + j) : (4 + 4 * mu + (j - 2));
                        h->fs[s].b[j] = bias[o];
                        for (int c = 0; c < kW; ++c)
                            for (int t = 0; t < 9; ++t)
                                h->fs[s].w[(c * 9 + t) * kActCh + j] = W[((size_t)o * kW + c) * 9 + t];
                    }
                }
            }
        }
    }
    // GELU table: handle-private device memory on the current device
    std::vector<double> tab(kGeluTableDoubles);
    build_gelu_table(tab.data());
    h->gelu_tab = nullptr;
    cudaError_t e = cudaGetDevice(&h->device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->gelu_tab, tab.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(h->gelu_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (h->gelu_tab) cudaFree(h->gelu_tab);
        delete h;
        return (int)e;
    }
    *out = h;
    return U1_OK;
}

int u1ft_destroy(u1ft_handle_t h) {
    if (!handle_ok(h)) return U1_ERR_HANDLE;
    h->magic = 0;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != h->device) cudaSetDevice(h->device);     // free on the device that owns the table / streams
    if (h->gelu_tab) cudaFree(h->gelu_tab);
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->lanes_ready) {
        cudaEventDestroy(h->ev_fork);
        for (int l = 0; l < kLanes; ++l) { cudaStreamDestroy(h->lane[l]); cudaEventDestroy(h->ev_join[l]); }
    }
    if (prev >= 0 && prev != h->device) cudaSetDevice(prev);
    delete h;
    return U1_OK;
}

size_t u1ft_workspace_bytes(u1ft_handle_t h, int B, int L, int with_backward) {
    if (!handle_ok(h) || B <= 0 || L <= 0) return 0;
    return (size_t)lanes_for(B, L) * carve_ft(nullptr, h->n_layers, chunk_chains(B, L), L, with_backward != 0).bytes;
}

int u1ft_forward(u1ft_handle_t h, const double* theta, double* theta_ori, double* logdet,
                 int B, int L, void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    g_graph_kernels = 0;
    g_graph_replays = 0;
    if (!handle_ok(h) || !handle_on_current_device(h)) return U1_ERR_HANDLE;
    if (!theta || B <= 0 || L <= 0) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    const int Cmax = chunk_chains(B, L);
    U1_TRY(check_ws_ft(ws, ws_bytes, carve_ft(nullptr, h->n_layers, Cmax, L, false).bytes));
    FtWs w = carve_ft(ws, h->n_layers, Cmax, L, false);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per = (size_t)2 * L * L;
    for (int b0 = 0; b0 < B; b0 += Cmax) {
        const int C = (B - b0 < Cmax) ? B - b0 : Cmax;
        U1_TRY(copy_async(w.th[0], theta + (size_t)b0 * per, (size_t)C * per * 8, st));
        U1_TRY(ft_forward_chunk(h, w, C, L, false, logdet != nullptr, st));
        if (theta_ori) U1_TRY(copy_async(theta_ori + (size_t)b0 * per, w.th[8], (size_t)C * per * 8, st));
        if (logdet) U1_TRY(copy_async(logdet + b0, w.logdet, (size_t)C * 8, st));
    }
    return U1_OK;
}

int u1ft_phase(u1ft_handle_t h, const double* theta, int index, double* delta,
               int B, int L, void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    g_graph_kernels = 0;
    g_graph_replays = 0;
    if (!handle_ok(h) || !handle_on_current_device(h)) return U1_ERR_HANDLE;
    if (!theta || !delta || B <= 0 || L <= 0 || index < 0 || index > 7) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    const int Cmax = chunk_chains(B, L);
    U1_TRY(check_ws_ft(ws, ws_bytes, carve_ft(nullptr, h->n_layers, Cmax, L, false).bytes));
    FtWs w = carve_ft(ws, h->n_layers, Cmax, L, false);
    cudaStream_t st = (cudaStream_t)stream;
    const int V = L * L, nblk = blocks_for(V, kUpdThreads);
    const size_t per = (size_t)2 * V;
    for (int b0 = 0; b0 < B; b0 += Cmax) {
        const int C = (B - b0 < Cmax) ? B - b0 : Cmax;
        const double* th = theta + (size_t)b0 * per;
        k_ft_features<<<blocks_for((size_t)C * V, 256), 256, 0, st>>>(th, w.X, index, C * V, L);
        count_launch();
        U1_CHECK_LAUNCH();
        U1_TRY(cnn_forward(h, w, index, C, L, false, st));
        k_ft_update<<<dim3(nblk, C), kUpdThreads, 0, st>>>(th, w.Kk[index], delta + (size_t)b0 * per, nullptr,
                                                           index, L, nblk, 1);
        count_launch();
        U1_CHECK_LAUNCH();
    }
    return U1_OK;
}

int u1ft_coefficients(u1ft_handle_t h, const double* theta, int index, double* K,
                      int B, int L, void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    g_graph_kernels = 0;
    g_graph_replays = 0;
    if (!handle_ok(h) || !handle_on_current_device(h)) return U1_ERR_HANDLE;
    if (!theta || !K || B <= 0 || L <= 0 || index < 0 || index > 7) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    const int Cmax = chunk_chains(B, L);
    U1_TRY(check_ws_ft(ws, ws_bytes, carve_ft(nullptr, h->n_layers, Cmax, L, false).bytes));
    FtWs w = carve_ft(ws, h->n_layers, Cmax, L, false);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t V = (size_t)L * L;
    for (int b0 = 0; b0 < B; b0 += Cmax) {
        const int C = (B - b0 < Cmax) ? B - b0 : Cmax;
        k_ft_features<<<blocks_for((size_t)C * V, 256), 256, 0, st>>>(theta + (size_t)b0 * 2 * V, w.X, index, (int)(C * V), L);
        count_launch();
        U1_CHECK_LAUNCH();
        FtWs wk = w;
        wk.Kk[index] = K + (size_t)b0 * kW * V;            // final layer writes straight into the caller's buffer
        U1_TRY(cnn_forward(h, wk, index, C, L, false, st));
    }
    return U1_OK;
}

size_t u1ft_inverse_workspace_bytes(int B, int L) {
    if (B <= 0 || L <= 0) return 0;
    return al((size_t)B * blocks_for((size_t)L * L, kUpdThreads) * 2 * sizeof(double));
}

int u1ft_inverse_iter(const double* theta_curr, const double* theta_iter, const double* K, int index,
                      double* theta_next, double* norms, int B, int L,
                      void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    g_graph_kernels = 0;
    g_graph_replays = 0;
    if (!theta_curr || !theta_iter || !K || !theta_next || !norms || B <= 0 || L <= 0 || index < 0 || index > 7)
        return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    if (theta_next == theta_iter || theta_next == theta_curr || B > 65535) return U1_ERR_ARG;
    U1_TRY(check_ws_ft(ws, ws_bytes, u1ft_inverse_workspace_bytes(B, L)));
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = blocks_for((size_t)L * L, kUpdThreads);
    k_ft_inverse_iter<<<dim3(nblk, B), kUpdThreads, 0, st>>>(theta_curr, theta_iter, K, theta_next, (double*)ws, index, L, nblk);
    count_launch();
    U1_CHECK_LAUNCH();
    k_sum_pairs<<<1, 256, 0, st>>>((const double*)ws, (size_t)B * nblk, norms);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1ft_action(u1ft_handle_t h, const double* theta, double* S_ft, double* theta_ori,
                double* logdet, int B, int L, double beta,
                void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    g_graph_kernels = 0;
    g_graph_replays = 0;
    if (!handle_ok(h) || !handle_on_current_device(h)) return U1_ERR_HANDLE;
    if (!theta || !S_ft || B <= 0 || L <= 0) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    const int Cmax = chunk_chains(B, L);
    U1_TRY(check_ws_ft(ws, ws_bytes, carve_ft(nullptr, h->n_layers, Cmax, L, false).bytes));
    FtWs w = carve_ft(ws, h->n_layers, Cmax, L, false);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per = (size_t)2 * L * L;
    for (int b0 = 0; b0 < B; b0 += Cmax) {
        const int C = (B - b0 < Cmax) ? B - b0 : Cmax;
        U1_TRY(copy_async(w.th[0], theta + (size_t)b0 * per, (size_t)C * per * 8, st));
        U1_TRY(ft_forward_chunk(h, w, C, L, false, true, st));
        U1_TRY(ft_action_chunk(w, C, L, beta, nullptr, w.sums_new, S_ft + b0, nullptr, st));
        if (theta_ori) U1_TRY(copy_async(theta_ori + (size_t)b0 * per, w.th[8], (size_t)C * per * 8, st));
        if (logdet) U1_TRY(copy_async(logdet + b0, w.logdet, (size_t)C * 8, st));
    }
    return U1_OK;
}

int u1ft_force(u1ft_handle_t h, const double* theta, double* force, double* S_ft,
               int B, int L, double beta, void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    g_graph_kernels = 0;
    g_graph_replays = 0;
    if (!handle_ok(h) || !handle_on_current_device(h)) return U1_ERR_HANDLE;
    if (!theta || !force || B <= 0 || L <= 0) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    const int Cmax = chunk_chains(B, L);
    U1_TRY(check_ws_ft(ws, ws_bytes, carve_ft(nullptr, h->n_layers, Cmax, L, true).bytes));
    FtWs w = carve_ft(ws, h->n_layers, Cmax, L, true);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per = (size_t)2 * L * L;
    for (int b0 = 0; b0 < B; b0 += Cmax) {
        const int C = (B - b0 < Cmax) ? B - b0 : Cmax;
        U1_TRY(copy_async(w.th[0], theta + (size_t)b0 * per, (size_t)C * per * 8, st));
        double* res = nullptr;
        U1_TRY(ft_force_chunk(h, w, C, L, beta, S_ft ? S_ft + b0 : nullptr, &res, st));
        U1_TRY(copy_async(force + (size_t)b0 * per, res, (size_t)C * per * 8, st));
    }
    return U1_OK;
}

// leapfrog of one chunk: state in (w.th[0], w.pi), both updated in place (hmc_u1_ft.py:154-179).
static int ft_leapfrog_chunk(u1ft_handle_t h, const FtWs& w, int C, int L, double beta, double dt,
                             int n_steps, cudaStream_t st) {
    const size_t n = (size_t)C * 2 * L * L;
    auto force_kick = [&](cudaStream_t s) -> int {   // pi -= dt * dS_FT/dtheta(theta): one graph replay per step
        double* F = nullptr;
        U1_TRY(ft_force_chunk(h, w, C, L, beta, nullptr, &F, s));
        k_axpy<<<ew_grid_ft(n), 256, 0, s>>>(w.pi, F, -dt, n);
        count_launch();
        U1_CHECK_LAUNCH();
        return U1_OK;
    };
    for (int s = 0; s < n_steps; ++s) {
        k_axpy<<<ew_grid_ft(n), 256, 0, st>>>(w.th[0], w.pi, s == 0 ? 0.5 * dt : dt, n);
        count_launch();
        U1_CHECK_LAUNCH();
        U1_TRY(run_graphed(h, GRAPH_FORCE_KICK, (void*)w.th[0], C, L, beta, dt, st, force_kick));
    }
    k_axpy_wrap<<<ew_grid_ft(n), 256, 0, st>>>(w.th[0], w.pi, 0.5 * dt, n);
    count_launch();
    U1_CHECK_LAUNCH();
    return U1_OK;
}

int u1ft_leapfrog(u1ft_handle_t h, const double* theta_in, const double* pi_in,
                  double* theta_out, double* pi_out, int B, int L, double beta, double dt,
                  int n_steps, void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    g_graph_kernels = 0;
    g_graph_replays = 0;
    if (!handle_ok(h) || !handle_on_current_device(h)) return U1_ERR_HANDLE;
    if (!theta_in || !pi_in || !theta_out || !pi_out || B <= 0 || L <= 0 || n_steps < 1) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    const int Cmax = chunk_chains(B, L);
    U1_TRY(check_ws_ft(ws, ws_bytes, carve_ft(nullptr, h->n_layers, Cmax, L, true).bytes));
    FtWs w = carve_ft(ws, h->n_layers, Cmax, L, true);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per = (size_t)2 * L * L;
    for (int b0 = 0; b0 < B; b0 += Cmax) {
        const int C = (B - b0 < Cmax) ? B - b0 : Cmax;
        U1_TRY(copy_async(w.th[0], theta_in + (size_t)b0 * per, (size_t)C * per * 8, st));
        U1_TRY(copy_async(w.pi, pi_in + (size_t)b0 * per, (size_t)C * per * 8, st));
        U1_TRY(ft_leapfrog_chunk(h, w, C, L, beta, dt, n_steps, st));
        U1_TRY(copy_async(theta_out + (size_t)b0 * per, w.th[0], (size_t)C * per * 8, st));
        U1_TRY(copy_async(pi_out + (size_t)b0 * per, w.pi, (size_t)C * per * 8, st));
    }
    return U1_OK;
}

int u1ft_trajectory(u1ft_handle_t h, double* theta, const double* pi_inject, const double* u_inject,
                    uint64_t seed, uint64_t chain0, uint64_t traj0, int n_traj,
                    int B, int L, double beta, double dt, int n_steps,
                    double* out_dH, int32_t* out_accept, double* out_H, double* out_plaq,
                    double* out_topo, double* theta_ori_out,
                    void* ws, size_t ws_bytes, u1_stream_t stream) {
    g_launches = 0;
    g_graph_kernels = 0;
    g_graph_replays = 0;
    if (!handle_ok(h) || !handle_on_current_device(h)) return U1_ERR_HANDLE;
    if (!theta || B <= 0 || L <= 0 || n_traj < 1 || n_steps < 1) return U1_ERR_ARG;
    if (L & 1) return U1_ERR_SHAPE;
    const int Cmax = chunk_chains(B, L);
    const int nlanes = lanes_for(B, L);
    const size_t lane_bytes = carve_ft(nullptr, h->n_layers, Cmax, L, true).bytes;
    U1_TRY(check_ws_ft(ws, ws_bytes, (size_t)nlanes * lane_bytes));
    cudaStream_t user = (cudaStream_t)stream;
    cudaStream_t lane_st[kLanes];
    for (int l = 0; l < kLanes; ++l) lane_st[l] = user;
    if (nlanes > 1) {                                   // fork: lanes start after everything queued on `stream`
        U1_TRY(ensure_lanes(h));
        cudaError_t e = cudaEventRecord(h->ev_fork, user);
        for (int l = 0; l < nlanes && e == cudaSuccess; ++l) {
            lane_st[l] = h->lane[l];
            e = cudaStreamWaitEvent(h->lane[l], h->ev_fork, 0);
        }
        if (e != cudaSuccess) return (int)e;
    }
    const int V = L * L;
    const size_t per = (size_t)2 * V;
    const double inv_vol = 1.0 / (double)V;
    // chunk-outer, trajectory-inner; chunk i runs on lane i % nlanes with that lane's workspace
    int chunk_idx = 0;
    for (int b0 = 0; b0 < B; b0 += Cmax, ++chunk_idx) {
        const int C = (B - b0 < Cmax) ? B - b0 : Cmax;
        const int ln = chunk_idx % nlanes;
        cudaStream_t st = lane_st[ln];
        FtWs w = carve_ft((char*)ws + (size_t)ln * lane_bytes, h->n_layers, Cmax, L, true);
        const size_t n = (size_t)C * per;
        double* th_cur = theta + (size_t)b0 * per;          // current (accepted) state, caller memory
        for (int t = 0; t < n_traj; ++t) {
            const uint64_t traj = traj0 + (uint64_t)t;
            const size_t o = (size_t)t * B + b0;
            // momenta (hmc_u1_ft.py:195)
            if (pi_inject) U1_TRY(copy_async(w.pi, pi_inject + ((size_t)t * B + b0) * per, n * 8, st));
            else {
                k_momenta_ft<<<blocks_for((size_t)C * V, 256), 256, 0, st>>>(w.pi, seed, chain0 + b0, traj, C * V, V);
                count_launch();
                U1_CHECK_LAUNCH();
            }
            // H_old (hmc_u1_ft.py:196-197)
            U1_TRY(copy_async(w.th[0], th_cur, n * 8, st));
            auto action_pass = [&](cudaStream_t s) -> int {   // theta in w.th[0], momenta in w.pi -> (w.sums_new, w.H_new, w.th[8])
                U1_TRY(ft_forward_chunk(h, w, C, L, false, true, s));
                return ft_action_chunk(w, C, L, beta, w.pi, w.sums_new, nullptr, w.H_new, s);
            };
            U1_TRY(run_graphed(h, GRAPH_ACTION, (void*)w.th[0], C, L, beta, 0.0, st, action_pass));
            U1_TRY(copy_async(w.sums_old, w.sums_new, (size_t)C * 3 * 8, st));
            U1_TRY(copy_async(w.H_old, w.H_new, (size_t)C * 8, st));
            if (theta_ori_out && t == n_traj - 1)            // F(theta_old), kept if the step is rejected
                U1_TRY(copy_async(theta_ori_out + (size_t)b0 * per, w.th[8], n * 8, st));
            // leapfrog (hmc_u1_ft.py:199)
            U1_TRY(ft_leapfrog_chunk(h, w, C, L, beta, dt, n_steps, st));
            U1_TRY(copy_async(w.th_new, w.th[0], n * 8, st));
            // H_new (hmc_u1_ft.py:200-201)
            U1_TRY(run_graphed(h, GRAPH_ACTION, (void*)w.th[0], C, L, beta, 0.0, st, action_pass));
            k_ft_metropolis<<<blocks_for(C, 128), 128, 0, st>>>(
                w.H_old, w.H_new, w.sums_old, w.sums_new, u_inject ? u_inject + o : nullptr, seed,
                chain0 + b0, traj, C, inv_vol, out_dH ? out_dH + o : nullptr,
                out_accept ? out_accept + o : nullptr, out_H ? out_H + o : nullptr,
                out_plaq ? out_plaq + o : nullptr, out_topo ? out_topo + o : nullptr, w.acc);
            count_launch();
            U1_CHECK_LAUNCH();
            k_select_chain<<<ew_grid_ft(n), 256, 0, st>>>(th_cur, w.th_new, w.acc, per, n);
            count_launch();
            U1_CHECK_LAUNCH();
            if (theta_ori_out && t == n_traj - 1) {
                k_select_chain<<<ew_grid_ft(n), 256, 0, st>>>(theta_ori_out + (size_t)b0 * per, w.th[8], w.acc, per, n);
                count_launch();
                U1_CHECK_LAUNCH();
            }
        }
    }
    if (nlanes > 1) {                                   // join: `stream` continues after both lanes
        for (int l = 0; l < nlanes; ++l) {
            cudaError_t e = cudaEventRecord(h->ev_join[l], h->lane[l]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(user, h->ev_join[l], 0);
            if (e != cudaSuccess) return (int)e;
        }
    }
    return U1_OK;
}

}  // extern "C"

#include "u1_ft_train.cuh"
