/* Accuracy check of vb_log_pos / vb_rcp_pos (csrc/vb_common.cuh) against glibc: gcc -O2 -ffp-contract=off -mfma tools/log_check.c -lm
   rcp.approx.ftz.f64 (MUFU.RCP64H, ~20 good bits) is emulated by a single-precision reciprocal. */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static inline double rcp_approx(double x) { return (double)(1.0f / (float)x); }
static inline double vb_rcp_pos(double x) {
    double r = rcp_approx(x);
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
static inline double vb_log_pos(double x) {
    int64_t b; memcpy(&b, &x, 8);
    int32_t hi = (int32_t)(b >> 32); uint32_t lo = (uint32_t)b;
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    const int big = hi >= 0x3ff6a09f;
    hi -= big ? 0x00100000 : 0;
    e += big;
    b = ((int64_t)hi << 32) | lo;
    double m; memcpy(&m, &b, 8);
    const double f = m - 1.0;
    const double s = f * vb_rcp_pos(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01), 6.666666666666735130e-01);
    const double R = t1 + t2;
    const double hfsq = 0.5 * f * f;
    const double k = (double)e;
    return k * 6.93147180369123816490e-01 - ((hfsq - fma(s, hfsq + R, k * 1.90821492927058770002e-10)) - f);
}
int main() {
    double maxulp = 0, worst = 0, maxr = 0; srand(2);
    for (long i = 0; i < 20000000; ++i) {
        double u = rand() / (double)RAND_MAX, v = rand() / (double)RAND_MAX;
        double x = (i % 2) ? exp((u - 0.5) * 1300.0) : 0.5 + 1.5 * v;
        double a = vb_log_pos(x), b = log(x);
        double ulp = fabs(a - b) / fabs(nextafter(b, INFINITY) - b);
        if (b != 0 && ulp > maxulp) { maxulp = ulp; worst = x; }
        double xr = (i % 2) ? exp((u - 0.5) * 120.0) : x; double ra = vb_rcp_pos(xr), rb = 1.0 / xr;
        double ur = fabs(ra - rb) / (nextafter(rb, INFINITY) - rb);
        if (ur > maxr) maxr = ur;
    }
    printf("log: max ulp %.3f at %.17g ; rcp: max ulp %.3f\n", maxulp, worst, maxr);
    printf("%g %g %g\n", vb_log_pos(1.0), vb_log_pos(2.0) - log(2.0), vb_log_pos(1e-300) - log(1e-300));
    return 0;
}
