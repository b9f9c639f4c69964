// Host check of csrc/u1_fastmath.cuh fast_sin against 80-bit sinl:  g++ -O2 -mfma tools/check_fast_sin.cpp
#include <cstdio>
#include <cmath>
#include <random>
#include "../hmc_ft_b200/csrc/u1_fastmath.cuh"
int main() {
    std::mt19937_64 g(1);
    double worst = 0, wx = 0, worst_lib = 0;
    for (double range : {3.2, 13.0, 40.0, 1000.0}) {
        std::uniform_real_distribution<double> d(-range, range);
        double w = 0;
        for (long i = 0; i < 10000000; ++i) {
            double x = d(g);
            long double ref = sinl((long double)x);
            double e = fabs((double)((long double)u1::fast_sin(x) - ref));
            double el = fabs((double)((long double)sin(x) - ref));
            if (e > w) { w = e; }
            if (e > worst) { worst = e; wx = x; }
            if (el > worst_lib) worst_lib = el;
        }
        printf("range %.1f: max abs err %.3e\n", range, w);
    }
    double edge[] = {0.0, -0.0, M_PI, -M_PI, M_PI / 2, 1e-300, 4 * M_PI, -4 * M_PI, 1.5707963267948966, 4.71238898038469};
    for (double x : edge) printf("  x=%.17g fast=%.17g lib=%.17g\n", x, u1::fast_sin(x), sin(x));
    printf("worst %.3e at x=%.17g ; libm worst %.3e\n", worst, wx, worst_lib);
    return worst < 2.5e-16 ? 0 : 1;
}
