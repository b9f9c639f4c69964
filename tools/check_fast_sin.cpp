// Host check of csrc/u1_fastmath.cuh against 80-bit libm:  g++ -O2 -mfma -std=c++17 tools/check_fast_sin.cpp
//   fast_sin / fast_cos / fast_sincos vs sinl/cosl, wrap_fast vs the reference's regularize expression,
//   winding vs that expression's floor, neg2_log vs logl (uniform, near 0, near 1).
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <random>
#include "../hmc_ft_b200/csrc/u1_fastmath.cuh"
static const u1::LogTabEntry kTab[u1::kLogTabN] = {
#include "../hmc_ft_b200/csrc/u1_logtab.inc"
};
static double wrap_ref(double t) {                    // utils.py:182-187
    const double pi = 3.14159265358979323846;
    double w = (t - pi) / (2 * pi);
    return 2 * pi * (w - floor(w) - 0.5);
}
int main(int argc, char** argv) {
    const long N = argc > 1 ? atol(argv[1]) : 10000000;      // samples per range
    std::mt19937_64 g(1);
    double worst = 0, wx = 0, worst_lib = 0, worst_c = 0, worst_sc = 0, worst_w = 0;
    long wind_bad = 0;
    for (double range : {3.2, 13.0, 40.0, 1000.0}) {
        std::uniform_real_distribution<double> d(-range, range);
        double w = 0, wc = 0;
        for (long i = 0; i < N; ++i) {
            double x = d(g);
            long double ref = sinl((long double)x), refc = cosl((long double)x);
            double e = fabs((double)((long double)u1::fast_sin(x) - ref));
            double ec = fabs((double)((long double)u1::fast_cos(x) - refc));
            double el = fabs((double)((long double)sin(x) - ref));
            double s2, c2;
            u1::fast_sincos(x, s2, c2);
            double esc = fmax(fabs((double)((long double)s2 - ref)), fabs((double)((long double)c2 - refc)));
            if (e > w) w = e;
            if (ec > wc) wc = ec;
            if (e > worst) { worst = e; wx = x; }
            if (ec > worst_c) worst_c = ec;
            if (esc > worst_sc) worst_sc = esc;
            if (el > worst_lib) worst_lib = el;
            const double wr = wrap_ref(x), wf = u1::wrap_fast(x);
            double dw = fabs(wr - wf);
            if (dw > 3.0) dw = fabs(dw - 2 * M_PI);            // opposite sides of the cut, same angle
            if (dw > worst_w) worst_w = dw;
            if (!(wf >= -M_PI && wf < M_PI)) { printf("wrap_fast out of range at %.17g: %.17g\n", x, wf); return 1; }
            const int n = u1::winding(x);
            const double nref = floor((x - M_PI) / (2 * M_PI)) + 1.0;
            if ((double)n != nref && fabs(fabs(wr) - M_PI) > 1e-9) ++wind_bad;
        }
        printf("range %.1f: sin max abs err %.3e  cos %.3e\n", range, w, wc);
    }
    double edge[] = {0.0, -0.0, M_PI, -M_PI, M_PI / 2, 1e-300, 4 * M_PI, -4 * M_PI, 1.5707963267948966, 4.71238898038469};
    for (double x : edge)
        printf("  x=%.17g fast=%.17g lib=%.17g  cos fast=%.17g lib=%.17g wrap=%.17g ref=%.17g\n", x, u1::fast_sin(x), sin(x),
               u1::fast_cos(x), cos(x), u1::wrap_fast(x), wrap_ref(x));
    printf("sin worst %.3e at x=%.17g ; cos worst %.3e ; sincos worst %.3e ; libm sin worst %.3e\n", worst, wx, worst_c, worst_sc, worst_lib);
    printf("wrap_fast vs regularize: max abs diff %.3e ; winding mismatches away from the cut: %ld\n", worst_w, wind_bad);

    // -2 ln u
    double worst_rel = 0, worst_abs = 0, worst_u = 0;
    auto chk = [&](double u) {
        const long double ref = -2.0L * logl((long double)u);
        const double got = u1::neg2_log(u, kTab);
        const double ea = fabs((double)((long double)got - ref));
        const double er = ref != 0 ? ea / (double)fabsl(ref) : ea;
        if (er > worst_rel) { worst_rel = er; worst_u = u; }
        if (ea > worst_abs) worst_abs = ea;
    };
    std::uniform_real_distribution<double> d01(0.0, 1.0);
    for (long i = 0; i < 2 * N; ++i) {
        const uint64_t m = g() >> 11;
        chk(((double)m + 0.5) * (1.0 / 9007199254740992.0));
    }
    for (long i = 0; i < N / 5; ++i) {
        chk(fmin(1.0, ldexp(d01(g) + 1e-3, -(int)(g() % 54))));   // small u (domain is (0,1])
        chk(1.0 - ldexp(d01(g) + 1e-3, -(int)(g() % 50) - 1));  // u -> 1
        chk(0.70710678118654752 + 1e-6 * (d01(g) - 0.5));       // around the mantissa split
        chk(1.0 + 0.0 * d01(g));
    }
    chk(ldexp(1.0, -54)); chk(1.0); chk(0.5); chk(0.9999999999999999);
    printf("neg2_log: max rel err %.3e (u=%.17g), max abs err %.3e\n", worst_rel, worst_u, worst_abs);
    const bool ok = worst < 2.5e-16 && worst_c < 3.5e-16 && worst_sc < 3.5e-16 && worst_w < 2e-13 && wind_bad == 0 &&
                    worst_rel < 6e-16;
    printf(ok ? "OK\n" : "FAIL\n");
    return ok ? 0 : 1;
}
