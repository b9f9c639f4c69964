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

}  // namespace u1
