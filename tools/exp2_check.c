// Host emulation of csrc/fastmath2.cuh (same constants, same operation order, C99 fma) against long-double libm.
// Build: gcc -O2 -std=gnu99 -ffp-contract=off tools/exp2_check.c -lm -o /tmp/exp2_check ; prints the max error in ulp.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define E2TAB 512
static double tab[E2TAB];
static const double C1 = 1.3538030870311431e-03, C2 = 9.163913992275266e-07, C3 = 4.135378350676714e-10,
                    C4 = 1.399621994296973e-13;
static int loint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)u; }
static int hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)(u >> 32); }
static double hilo(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x; }
static double core(double t, double r) {
    const int N = loint(t);
    double q = C4;
    q = fma(q, r, C3); q = fma(q, r, C2); q = fma(q, r, C1);
    const double e = q * r;
    const double T = tab[N & (E2TAB - 1)];
    const double v = fma(T, e, T);
    return hilo(hiint(v) + ((N >> 9) << 20), loint(v));
}
static double pow2_fast(double p, double l2s) {
    const double MAGIC = 6755399441055744.0;
    const double t = fma(p, l2s, MAGIC), nf = t - MAGIC, r = fma(p, l2s, -nf);
    return core(t, r);
}
static double exp2_neg(double s) {
    const double MAGIC = 6755399441055744.0;
    const double t = fma(s, -512.0, MAGIC), nf = t - MAGIC, r = fma(s, -512.0, -nf);
    double v = core(t, r);
    if (s > 1021.0) v = 0.0;
    if (s < -1023.0) v = INFINITY;
    return v;
}
static double ulp_err(double got, long double ref) {
    if (ref == 0.0L) return got == 0.0 ? 0.0 : 1e9;
    int e; frexpl(ref, &e);
    return (double)fabsl(((long double)got - ref) / ldexpl(1.0L, e - 53));
}
int main(void) {
    for (int j = 0; j < E2TAB; ++j) tab[j] = (double)exp2l((long double)j / E2TAB);
    srand48(1);
    double worst_p = 0, worst_e = 0;
    for (int it = 0; it < 4000000; ++it) {
        const double dx = exp(log(1e-7) + drand48() * (log(30.0) - log(1e-7)));
        const double p = 0.25 + drand48() * 2.05;
        double l2s = 512.0 * log2(dx);  // the table value (rounded) is the kernel's input: measure against it
        const long double ref = powl(2.0L, (long double)p * (long double)l2s / 512.0L);
        const double e1 = ulp_err(pow2_fast(p, l2s), ref);
        if (e1 > worst_p) worst_p = e1;
        const double s = (it & 1) ? drand48() * 40.0 : exp(log(1e-9) + drand48() * log(1e12));
        const double e2 = ulp_err(exp2_neg(s), exp2l(-(long double)s));
        if (s < 1000 && e2 > worst_e) worst_e = e2;
    }
    // edges
    const double z = pow2_fast(1.0, -430.0 * 512.0), big = pow2_fast(2.3, 430.0 * 512.0);
    printf("{\"pow2_fast_max_ulp\": %.4f, \"exp2_neg_max_ulp\": %.4f, \"zero_dist\": %.3g, \"padding\": %.3g, \"exp2_neg_0\": %.17g, \"exp2_neg_1100\": %g, \"exp2_neg_inf\": %g}\n",
           worst_p, worst_e, z, big, exp2_neg(0.0), exp2_neg(1100.0), exp2_neg(INFINITY));
    return 0;
}
