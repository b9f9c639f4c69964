// Branch-free fp64 sin for bounded arguments (|x| < 2^20), used in the leapfrog inner loop where the
// argument is a plaquette angle.  ~14 FP64 instructions + 2 integer ones, everything in registers.
//   k = rint(x/pi) (magic-number rounding);  r = x - k*pi (two-term Cody-Waite, FMA);
//   sin x = (-1)^k sin r,  sin r = r + r*u*q(u), u = r^2, q = degree-7 near-minimax polynomial on
//   [0,(pi/2+0.02)^2] (Chebyshev fit of (sin(sqrt u)/sqrt u - 1)/u, approximation error 1.6e-18).
// Measured against 80-bit sinl on 4e7 arguments in [-1000,1000]: max abs error 2.2e-16 (tools/check_fast_sin.cpp).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define U1_HD __host__ __device__ __forceinline__
#else
#define U1_HD inline
#endif

namespace u1 {

U1_HD double fast_sin(double x) {
    const double kInvPi = 0x1.45f306dc9c883p-2;
    const double kMagic = 6755399441055744.0;           // 1.5 * 2^52
    const double kPiHi = 0x1.921fb54442d18p+1, kPiLo = 0x1.1a62633145c07p-53;
    const double t = fma(x, kInvPi, kMagic);
    const double kd = t - kMagic;
    double r = fma(-kd, kPiHi, x);
    r = fma(-kd, kPiLo, r);
#if defined(__CUDA_ARCH__)
    const int k = __double2loint(t);
    r = __hiloint2double(__double2hiint(r) ^ (k << 31), __double2loint(r));
#else
    uint64_t tb, rb;
    memcpy(&tb, &t, 8);
    memcpy(&rb, &r, 8);
    rb ^= (tb & 1ull) << 63;
    memcpy(&r, &rb, 8);
#endif
    const double u = r * r;
    double q = 0x1.8959f439ca300p-49;
    q = fma(q, u, -0x1.ae4edb089e6c0p-41);
    q = fma(q, u, 0x1.6123f9d9c565ep-33);
    q = fma(q, u, -0x1.ae64558a62d07p-26);
    q = fma(q, u, 0x1.71de3a543121ap-19);
    q = fma(q, u, -0x1.a01a01a01872dp-13);
    q = fma(q, u, 0x1.1111111111105p-7);
    q = fma(q, u, -0x1.5555555555555p-3);
    return fma(r * u, q, r);
}


// Shared argument reduction of fast_sin / fast_cos / fast_sincos: x = k*pi + r, |r| <= pi/2 (+ rounding);
// returns r and the parity of k in the sign-bit position (bit 31 of the high word).
U1_HD double reduce_pi(double x, unsigned& sign_hi) {
    const double kInvPi = 0x1.45f306dc9c883p-2;
    const double kMagic = 6755399441055744.0;           // 1.5 * 2^52
    const double kPiHi = 0x1.921fb54442d18p+1, kPiLo = 0x1.1a62633145c07p-53;
    const double t = fma(x, kInvPi, kMagic);
    const double kd = t - kMagic;
    double r = fma(-kd, kPiHi, x);
    r = fma(-kd, kPiLo, r);
#if defined(__CUDA_ARCH__)
    sign_hi = (unsigned)__double2loint(t) << 31;
#else
    uint64_t tb;
    memcpy(&tb, &t, 8);
    sign_hi = (unsigned)(tb & 1ull) << 31;
#endif
    return r;
}
U1_HD double flip_sign(double v, unsigned sign_hi) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (int)sign_hi, __double2loint(v));
#else
    uint64_t b;
    memcpy(&b, &v, 8);
    b ^= (uint64_t)sign_hi << 32;
    memcpy(&v, &b, 8);
    return v;
#endif
}
// sin r for |r| <= pi/2 + 0.02 (the polynomial of fast_sin).
U1_HD double sin_poly(double r, double u) {
    double q = 0x1.8959f439ca300p-49;
    q = fma(q, u, -0x1.ae4edb089e6c0p-41);
    q = fma(q, u, 0x1.6123f9d9c565ep-33);
    q = fma(q, u, -0x1.ae64558a62d07p-26);
    q = fma(q, u, 0x1.71de3a543121ap-19);
    q = fma(q, u, -0x1.a01a01a01872dp-13);
    q = fma(q, u, 0x1.1111111111105p-7);
    q = fma(q, u, -0x1.5555555555555p-3);
    return fma(r * u, q, r);
}
// cos r = 1 + u q(u), u = r^2: degree-7 Chebyshev fit of (cos(sqrt u) - 1)/u on [0,(pi/2+0.02)^2]
// (approximation error 7.8e-18; mpmath, 200 bits).
U1_HD double cos_poly(double u) {
    double q = 0x1.a07aa75757565p-45;
    q = fma(q, u, -0x1.935df1dda4f80p-37);
    q = fma(q, u, 0x1.1eed1482feb0ep-29);
    q = fma(q, u, -0x1.27e4fa555c0c2p-22);
    q = fma(q, u, 0x1.a01a019d2ef71p-16);
    q = fma(q, u, -0x1.6c16c16c1310bp-10);
    q = fma(q, u, 0x1.555555555551cp-5);
    q = fma(q, u, -0x1.0000000000000p-1);
    return fma(u, q, 1.0);
}
// Branch-free cos for bounded arguments, 13 FP64 instructions; max abs error 2.3e-16 (tools/check_fast_sin.cpp).
U1_HD double fast_cos(double x) {
    unsigned sg;
    const double r = reduce_pi(x, sg);
    return flip_sign(cos_poly(r * r), sg);
}
U1_HD void fast_sincos(double x, double& s, double& c) {
    unsigned sg;
    const double r = reduce_pi(x, sg);
    const double u = r * r;
    s = flip_sign(sin_poly(r, u), sg);
    c = flip_sign(cos_poly(u), sg);
}

// t - 2 pi rint(t / 2 pi), folded into [-pi, pi) like utils.py:182-187 regularize (whose own
// expression, wrap_pi() in u1_common.cuh, costs a division and a floor): 4 FP64 instructions plus the
// two range fix-ups that make the interval half-open.  Agrees with wrap_pi to its rounding (~|t| 2^-52).
U1_HD double wrap_fast(double t) {
    const double kInv2Pi = 0x1.45f306dc9c883p-3;
    const double kMagic = 6755399441055744.0;
    const double k2PiHi = 0x1.921fb54442d18p+2, k2PiLo = 0x1.1a62633145c07p-52;
    const double kPi = 0x1.921fb54442d18p+1;
    const double kd = fma(t, kInv2Pi, kMagic) - kMagic;
    double r = fma(-kd, k2PiHi, t);
    r = fma(-kd, k2PiLo, r);
    if (r >= kPi) r -= k2PiHi;
    if (r < -kPi) r += k2PiHi;
    return r;
}
// Winding number n = rint(t / 2 pi) of an angle (so that t - 2 pi n is the wrapped angle): the
// topological charge of utils.py:189-196 is -sum_sites n(P) because sum_sites P = 0 identically.
U1_HD int winding(double t) {
    const double kInv2Pi = 0x1.45f306dc9c883p-3;
    const double kMagic = 6755399441055744.0;
    const double m = fma(t, kInv2Pi, kMagic);
#if defined(__CUDA_ARCH__)
    return __double2loint(m);
#else
    uint64_t b;
    memcpy(&b, &m, 8);
    return (int)(uint32_t)b;
#endif
}

// -2 ln(u) for u in (0,1]: u = 2^e m, m in [sqrt(1/2), sqrt(2)); interval i = top 7 bits of the
// re-biased high word of m; ln m = log1p(m x_i - 1) + y_i with the table {x_i = fl(1/c_i), y_i = -ln x_i}
// of u1_logtab.inc (tools/gen_log_table.py).  |m x_i - 1| < 3.9e-3 so the degree-6 Taylor polynomial of
// log1p truncates below 2e-18 (relative); the interval around 1 uses c = 1, so u -> 1 keeps full
// relative accuracy.  10 FP64 instructions + one int->double conversion.
struct LogTabEntry { double x, y; };
constexpr int kLogTabN = 128;
U1_HD double neg2_log(double u, const LogTabEntry* tab) {
    uint32_t hi, lo;
#if defined(__CUDA_ARCH__)
    hi = (uint32_t)__double2hiint(u);
    lo = (uint32_t)__double2loint(u);
#else
    uint64_t ub;
    memcpy(&ub, &u, 8);
    hi = (uint32_t)(ub >> 32);
    lo = (uint32_t)ub;
#endif
    const uint32_t h2 = hi + (0x3ff00000u - 0x3fe6a09eu);
    const int e = (int)(h2 >> 20) - 0x3ff;
    const uint32_t frac = h2 & 0x000fffffu;
    const uint32_t h3 = frac + 0x3fe6a09eu;
    const LogTabEntry te = tab[frac >> 13];
    double m;
#if defined(__CUDA_ARCH__)
    m = __hiloint2double((int)h3, (int)lo);
#else
    ub = ((uint64_t)h3 << 32) | lo;
    memcpy(&m, &ub, 8);
#endif
    const double r = fma(m, te.x, -1.0);
    double q = -1.0 / 6.0;
    q = fma(q, r, 0.2);
    q = fma(q, r, -0.25);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -0.5);
    const double p = fma(r * r, q, r);
    const double s = te.y + p;
    return fma(-2.0, s, (double)e * -0x1.62e42fefa39efp+0);   // -2 (e ln2 + s)
}

// sqrt(w) for finite w > 0 without the special-case paths of sqrt(): MUFU.RSQ64H seed (~2^-21), one
// Goldschmidt iteration and a final Newton correction (error ~2^-80 before the last rounding).
U1_HD double sqrt_pos(double w) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w));
    double g = w * y, h = 0.5 * y;
    const double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    const double d = fma(-g, g, w);
    return fma(d, h, g);
#else
    return sqrt(w);
#endif
}

}  // namespace u1
