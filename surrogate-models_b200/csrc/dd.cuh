// Double-double (~106-bit) arithmetic and correctly-rounded exp / pow for the extended-precision evaluation of
// ill-conditioned candidates (krig_rescue.cu).  Plain C++ with explicit fused multiply-adds: the same source builds
// for the device (nvcc) and for the host (g++ -ffp-contract=off, tools/dd_check.cpp), so the arithmetic can be checked
// against glibc / libquadmath in the build container without a GPU.
//
// Why this exists (reference: mylibs/kriging.py:70-77,95-116): the reference forms R_ij = exp(-sum_d theta_d |dx|^p_d) with
// libm's pow/exp (glibc: correctly rounded in all but ~1e-5 of the calls) and factorises that fp64 matrix.  Where
// cond(R) approaches 1/eps, whether R "is" positive definite -- and therefore which candidate wins a grid search --
// depends on the last bit of its entries and on the rounding of every pivot.  The rescue path therefore (a) reproduces
// the reference's entries bit for bit (correctly-rounded pow and exp, fp64 multiply/add in the reference's order) and
// (b) factorises that matrix in double-double, i.e. returns the exact likelihood of the matrix the reference holds.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define DD_FN __host__ __device__ __forceinline__
#else
#define DD_FN inline
#endif

namespace smb {

#if defined(__CUDA_ARCH__)
#define DD_ADD(a, b) __dadd_rn((a), (b))
#define DD_MUL(a, b) __dmul_rn((a), (b))
#define DD_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
// host build: compile with -ffp-contract=off so these stay separate roundings
#define DD_ADD(a, b) ((a) + (b))
#define DD_MUL(a, b) ((a) * (b))
#define DD_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

struct dd {
    double hi, lo;
};

DD_FN dd dd_make(double h, double l = 0.0) { return dd{h, l}; }
DD_FN dd two_sum(double a, double b) {
    const double s = DD_ADD(a, b), bb = DD_ADD(s, -a);
    return dd{s, DD_ADD(DD_ADD(a, -DD_ADD(s, -bb)), DD_ADD(b, -bb))};
}
DD_FN dd quick_two_sum(double a, double b) {  // |a| >= |b|
    const double s = DD_ADD(a, b);
    return dd{s, DD_ADD(b, -DD_ADD(s, -a))};
}
DD_FN dd two_prod(double a, double b) {
    const double p = DD_MUL(a, b);
    return dd{p, DD_FMA(a, b, -p)};
}
DD_FN dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
DD_FN dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo = DD_ADD(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = DD_ADD(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
DD_FN dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
DD_FN dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo = DD_ADD(s.lo, a.lo);
    return quick_two_sum(s.hi, s.lo);
}
DD_FN dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = DD_FMA(a.hi, b.lo, p.lo);
    p.lo = DD_FMA(a.lo, b.hi, p.lo);
    return quick_two_sum(p.hi, p.lo);
}
DD_FN dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo = DD_FMA(a.lo, b, p.lo);
    return quick_two_sum(p.hi, p.lo);
}
DD_FN dd dd_sqr(dd a) {
    dd p = two_prod(a.hi, a.hi);
    p.lo = DD_FMA(DD_MUL(2.0, a.hi), a.lo, p.lo);
    return quick_two_sum(p.hi, p.lo);
}
// c - a * b, the multiply-subtract of the factorisation
DD_FN dd dd_fnma(dd a, dd b, dd c) { return dd_sub(c, dd_mul(a, b)); }
DD_FN dd dd_div(dd a, dd b) {
    const double q1 = a.hi / b.hi;
    dd r = dd_sub(a, dd_mul_d(b, q1));
    const double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul_d(b, q2));
    const double q3 = r.hi / b.hi;
    dd q = quick_two_sum(q1, q2);
    return dd_add_d(q, q3);
}
DD_FN dd dd_recip(dd b) { return dd_div(dd{1.0, 0.0}, b); }
DD_FN dd dd_sqrt(dd a) {  // a > 0 (Karp / Markstein: one correction of the fp64 root)
    const double x = 1.0 / sqrt(a.hi);
    const double ax = DD_MUL(a.hi, x);
    const dd err = dd_sub(a, two_prod(ax, ax));
    return dd_add_d(two_sum(ax, DD_MUL(err.hi, DD_MUL(x, 0.5))), 0.0);
}

DD_FN double dd_ldexp(double v, int e) {  // exact scaling by 2^e, |e| <= 2200, without library calls in the hot path
#if defined(__CUDA_ARCH__)
    return ldexp(v, e);
#else
    return std::ldexp(v, e);
#endif
}

// exp(x) to ~2^-95 relative: x = k ln2 + r, exp(r) = (exp(r / 512))^512 with a 10-term Taylor series of exp(m) - 1 and
// nine squarings in the form q <- 2q + q^2 (keeps the leading 1 out of the sum).
DD_FN dd exp_dd(dd x) {
    if (!(x.hi > -745.2)) return dd{x.hi != x.hi ? x.hi : 0.0, 0.0};
    if (x.hi > 709.7) return dd{INFINITY, 0.0};
    const double kf = rint(DD_MUL(x.hi, 1.4426950408889634));
    const dd ln2 = dd{0.6931471805599453, 2.3190468138462996e-17};
    dd r = dd_sub(x, dd_mul_d(ln2, kf));
    r.hi = DD_MUL(r.hi, 0.001953125);  // / 512, exact
    r.lo = DD_MUL(r.lo, 0.001953125);
    // 1/j!, j = 10 .. 3, as fp64 (their terms are < 1e-10 of the sum); 1/2 exact
    const double c[8] = {2.755731922398589e-07, 2.7557319223985893e-06, 2.48015873015873e-05, 0.0001984126984126984,
                         0.001388888888888889,  0.008333333333333333,   0.041666666666666664, 0.16666666666666666};
    dd q = dd{c[0], 0.0};
#pragma unroll
    for (int j = 1; j < 8; ++j) q = dd_add_d(dd_mul(q, r), c[j]);
    q = dd_add_d(dd_mul(q, r), 0.5);
    q = dd_add_d(dd_mul(q, r), 1.0);
    q = dd_mul(q, r);  // exp(m) - 1, |m| <= 6.8e-4; absolute error < 1e-26 (the fp64 coefficients contribute < 4e-27)
#pragma unroll
    for (int s = 0; s < 9; ++s) q = dd_add(dd_mul_d(q, 2.0), dd_sqr(q));
    dd e = dd_add_d(q, 1.0);
    const int k = (int)kf;
    e.hi = dd_ldexp(e.hi, k);
    e.lo = dd_ldexp(e.lo, k);
    return e;
}

// ln(x), x > 0 finite: one Newton step on the fp64 logarithm, y = y0 + (x exp(-y0) - 1) - (..)^2 / 2
DD_FN dd log_dd(double x) {
    const double y0 = log(x);
    const dd z = dd_mul_d(exp_dd(dd{-y0, 0.0}), x);
    const dd dl = dd_add_d(z, -1.0);
    dd y = dd_add(dd{y0, 0.0}, dl);
    return dd_add_d(y, -0.5 * dl.hi * dl.hi);
}

// correctly rounded exp(x) (up to the ~2^-40 fraction of arguments whose value lies within 2^-95 of a rounding boundary)
DD_FN double exp_cr(double x) { return exp_dd(dd{x, 0.0}).hi; }

// correctly rounded pow(x, p) for x >= 0 (the reference's abs(dx) ** p, mylibs/kriging.py:75)
DD_FN double pow_cr(double x, double p) {
    if (p == 0.0) return 1.0;
    if (x == 0.0) return p > 0.0 ? 0.0 : INFINITY;
    if (p == 1.0) return x;
    if (p == 2.0) return DD_MUL(x, x);
    if (x == 1.0) return 1.0;
    return exp_dd(dd_mul_d(log_dd(x), p)).hi;
}

}  // namespace smb
