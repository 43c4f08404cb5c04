/* Accuracy check of vb_exp_nonpos (csrc/vb_common.cuh) against glibc exp: gcc -O2 -ffp-contract=off -mfma tools/exp_check.c -lm */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static inline double vb_exp_nonpos(double x) {
    const double SHIFT = 6755399441055744.0;
    x = x < -1000.0 ? -1000.0 : x;
    const double t = fma(x, 1.4426950408889634, SHIFT);
    int64_t tb; memcpy(&tb, &t, 8);
    const int n = (int)(uint32_t)tb;
    const double nf = t - SHIFT;
    double r = fma(nf, -6.93147180369123816490e-01, x);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;            // 1/13!
    p = fma(p, r, 2.08767569878681e-09);          // 1/12!
    p = fma(p, r, 2.505210838544172e-08);         // 1/11!
    p = fma(p, r, 2.755731922398589e-07);         // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);        // 1/9!
    p = fma(p, r, 2.48015873015873e-05);          // 1/8!
    p = fma(p, r, 0.0001984126984126984);         // 1/7!
    p = fma(p, r, 0.001388888888888889);          // 1/6!
    p = fma(p, r, 0.008333333333333333);          // 1/5!
    p = fma(p, r, 0.041666666666666664);          // 1/4!
    p = fma(p, r, 0.16666666666666666);           // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int64_t pb; memcpy(&pb, &p, 8);
    pb += (int64_t)n << 52;
    double out; memcpy(&out, &pb, 8);
    return n < -1000 ? 0.0 : out;
}
int main() {
    double maxulp = 0, worst = 0; srand(1);
    for (long i = 0; i < 20000000; ++i) {
        double u = rand() / (double)RAND_MAX;
        double x = (i % 3 == 0) ? -u * 690 : (i % 3 == 1 ? -u * 40 : -u * u * u);
        double a = vb_exp_nonpos(x), b = exp(x);
        double ulp = fabs(a - b) / (nextafter(b, INFINITY) - b);
        if (ulp > maxulp) { maxulp = ulp; worst = x; }
    }
    printf("max ulp %.3f at %.17g\n", maxulp, worst);
    printf("%g %g %g %g\n", vb_exp_nonpos(0.0), vb_exp_nonpos(-1e300), vb_exp_nonpos(-690.0) / exp(-690.0), vb_exp_nonpos(NAN));
    for (double x = -1.0; x > -1e301; x *= 1.7) if (vb_exp_nonpos(x) > exp(x) * 1.0000001 || (x > -690 && vb_exp_nonpos(x) < exp(x) * 0.9999999)) { printf("BAD at %g: %g vs %g\n", x, vb_exp_nonpos(x), exp(x)); return 1; }
    return 0;
}
