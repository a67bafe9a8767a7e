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
// "sloppy" sum (11 instead of 20 operations): the error is bounded relative to |a| + |b| instead of |a + b|.  That is
// exactly what a factorisation update c - l l' needs -- the entries are O(1) and an absolute error of ~1e-32 is 16 orders
// below the smallest pivot this path ever sees -- so the multiply-subtract of the inner loop uses it.
DD_FN dd dd_add_fast(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    s.lo = DD_ADD(s.lo, DD_ADD(a.lo, b.lo));
    return quick_two_sum(s.hi, s.lo);
}
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
DD_FN dd dd_fnma(dd a, dd b, dd c) { return dd_add_fast(c, dd_neg(dd_mul(a, b))); }
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

// 2^(j/64), j = 0..63, as (hi, lo) pairs rounded from 250-bit values
#define DD_EXP2_64_TABLE { \
        {1.0, 0.0}, {1.0108892860517005, -1.5234778603368577e-17}, \
        {1.0218971486541166, 5.109225028973444e-17}, {1.0330248790212284, 7.600838874027088e-18}, \
        {1.0442737824274138, 8.551889705537965e-17}, {1.0556451783605572, 1.759325738772092e-18}, \
        {1.0671404006768237, -7.899853966841582e-17}, {1.0787607977571199, -6.656660436056593e-17}, \
        {1.0905077326652577, -3.046782079812471e-17}, {1.102382583307841, 5.2660368715706944e-17}, \
        {1.1143867425958924, 1.0410278456845571e-16}, {1.1265216186082418, 5.165856758795457e-17}, \
        {1.1387886347566916, 8.912812676025408e-17}, {1.1511892299529827, 3.250710218863827e-17}, \
        {1.1637248587775775, 3.8292048369240935e-17}, {1.1763969916502812, 5.554203254218079e-17}, \
        {1.189207115002721, 3.982015231465646e-17}, {1.202156731452703, 6.644981499252301e-17}, \
        {1.215247359980469, -7.712630692681488e-17}, {1.22848053610687, -1.89878163130253e-17}, \
        {1.241857812073484, 4.658027591836937e-17}, {1.255380757024691, -6.7113898212968784e-18}, \
        {1.2690509571917332, 2.667932131342186e-18}, {1.2828700160787783, 1.713594918243561e-17}, \
        {1.2968395546510096, 2.5382502794888315e-17}, {1.3109612115247644, -7.181536135519454e-17}, \
        {1.3252366431597413, -2.8587312100388614e-17}, {1.339667524053303, 8.927282594831732e-17}, \
        {1.3542555469368927, 7.70094837980299e-17}, {1.3690024229745905, 9.593797919118849e-17}, \
        {1.383909881963832, -6.770511658794786e-17}, {1.3989796725383112, -9.614213209051323e-17}, \
        {1.4142135623730951, -9.667293313452913e-17}, {1.42961333839197, -1.2031642489053655e-17}, \
        {1.4451808069770467, -3.0237581349939873e-17}, {1.460917794180647, -5.600377186075216e-17}, \
        {1.4768261459394993, -3.483994556892796e-17}, {1.4929077282912648, 1.4192920154284036e-17}, \
        {1.5091644275934228, -1.016455327754295e-16}, {1.5255981507445384, -1.1024941712342561e-16}, \
        {1.5422108254079407, 7.949834809697621e-17}, {1.559004400237837, 3.7812070533575275e-17}, \
        {1.5759808451078865, -1.0136916471278304e-17}, {1.593142151342267, -1.0094406542311964e-16}, \
        {1.6104903319492543, 2.4707192569797888e-17}, {1.6280274218573478, -6.712955084707084e-17}, \
        {1.645755478153965, -1.0125679913674773e-16}, {1.6636765803267364, 5.8909926967131e-17}, \
        {1.681792830507429, 8.199010020581497e-17}, {1.7001063537185235, -8.0237193703977e-18}, \
        {1.718619298122478, -1.851380418263111e-17}, {1.7373338352737062, 3.164389299292957e-17}, \
        {1.7562521603732995, 2.960140695448873e-17}, {1.7753764925265212, 6.429731796556572e-17}, \
        {1.7947090750031072, 1.8227458427912087e-17}, {1.8142521755003989, -9.969531538920349e-17}, \
        {1.8340080864093424, 3.283107224245627e-17}, {1.8539791250833855, 9.761887490727594e-17}, \
        {1.8741676341103, -6.122763413004143e-17}, {1.8945759815869656, 3.4034035352165297e-17}, \
        {1.9152065613971474, -1.0619946056195963e-16}, {1.9360617934922943, 1.0332385960676326e-16}, \
        {1.9571441241754002, 8.960767791036668e-17}, {1.978456026387951, 4.0388753109278167e-17}}
static const double dd_exp2_64_host[64][2] = DD_EXP2_64_TABLE;
#if defined(__CUDACC__)
static __device__ const double dd_exp2_64_dev[64][2] = DD_EXP2_64_TABLE;
#endif

// exp(x) to ~2^-85 relative: x = (64 m + j) ln2/64 + r, |r| <= ln2/128, exp(x) = 2^m * 2^(j/64) * exp(r).  2^(j/64) comes from
// a 64-entry double-double table; exp(r) - 1 = r + r^2 (1/2 + r/6) + r^4 (1/24 + ... + r^7/11!) with the r^4 tail in fp64 (it
// is < 3.6e-11, so its rounding error is < 1e-26) and everything else in double-double.
DD_FN dd exp_dd(dd x) {
    if (!(x.hi > -745.2)) return dd{x.hi != x.hi ? x.hi : 0.0, 0.0};
    if (x.hi > 709.7) return dd{INFINITY, 0.0};
#if defined(__CUDA_ARCH__)
    const double(*EXP2_64)[2] = dd_exp2_64_dev;
#else
    const double(*EXP2_64)[2] = dd_exp2_64_host;
#endif

    const double kf = rint(DD_MUL(x.hi, 92.33248261689366));   // 64 / ln 2
    const dd ln2_64 = dd{0.010830424696249145, 3.623510646634843e-19};
    const dd r = dd_sub(x, dd_mul_d(ln2_64, kf));
    const double rh = r.hi;
    double t = 2.505210838544172e-08;                                     // 1/11!
    t = DD_FMA(t, rh, 2.755731922398589e-07);                             // 1/10!
    t = DD_FMA(t, rh, 2.7557319223985893e-06);
    t = DD_FMA(t, rh, 2.48015873015873e-05);
    t = DD_FMA(t, rh, 0.0001984126984126984);
    t = DD_FMA(t, rh, 0.001388888888888889);
    t = DD_FMA(t, rh, 0.008333333333333333);
    t = DD_FMA(t, rh, 0.041666666666666664);                              // 1/4!
    const double r2h = DD_MUL(rh, rh);
    t = DD_MUL(t, DD_MUL(r2h, r2h));
    const dd s = dd_add_d(dd_mul(r, dd{0.16666666666666666, 9.25185853854297e-18}), 0.5);
    dd q = dd_add(r, dd_mul(dd_sqr(r), s));
    q = dd_add_d(q, t);                                                   // exp(r) - 1
    const int k = (int)kf;
    const int j = k & 63, m = (k - j) / 64;
    const dd tj = dd{EXP2_64[j][0], EXP2_64[j][1]};
    dd e = dd_add(tj, dd_mul(tj, q));
    e.hi = dd_ldexp(e.hi, m);
    e.lo = dd_ldexp(e.lo, m);
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
