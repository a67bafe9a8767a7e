// Host build of csrc/dd.cuh for checking the extended-precision ("rescue") arithmetic in the build container.
// Serial restatement of krig_rescue.cu's per-candidate algorithm with the SAME dd primitives, exported through a
// C interface for oracle/dd_check.py (ctypes).  Test tooling only.
//   g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC tools/dd_check.cpp -o /tmp/libdd_check.so
#include <cmath>
#include <cstdlib>
#include <vector>
#include "../surrogate-models_b200/csrc/dd.cuh"
using namespace smb;

extern "C" {
void exp_cr_vec(const double* x, long n, double* out) {
    for (long i = 0; i < n; ++i) out[i] = exp_cr(x[i]);
}
void pow_cr_vec(const double* x, const double* p, long n, double* out) {
    for (long i = 0; i < n; ++i) out[i] = pow_cr(x[i], p[i]);
}
// entries exactly as the reference composes them (kriging.py:70-77) with correctly rounded pow / exp
void cr_entries(const double* X, int n, int k, const double* theta, const double* p, double* R) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
            for (int d = 0; d < k; ++d) s = s + theta[d] * pow_cr(fabs(X[i * k + d] - X[j * k + d]), p[d]);
            R[i * n + j] = R[j * n + i] = exp_cr(-s);
        }
}
static double dd_nll_one(const double* X, const double* y, int n, int k, const double* theta, const double* p, double* minpiv) {
    const int m = n + 2;
    std::vector<dd> A((size_t)m * (m + 1) / 2);
    auto at = [&](int i, int j) -> dd& { return A[(size_t)i * (i + 1) / 2 + j]; };
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
            for (int d = 0; d < k; ++d) s = s + theta[d] * pow_cr(fabs(X[i * k + d] - X[j * k + d]), p[d]);
            at(i, j) = dd{exp_cr(-s), 0.0};
        }
    for (int j = 0; j < n; ++j) { at(n, j) = dd{y[j], 0.0}; at(n + 1, j) = dd{1.0, 0.0}; }
    double mp = 1e300, logdet = 0.0;
    for (int j = 0; j < n; ++j) {
        const dd d = at(j, j);
        if (d.hi < mp) mp = d.hi;
        if (!(d.hi > 0.0)) { *minpiv = d.hi; return INFINITY; }
        logdet += log(d.hi) + d.lo / d.hi;
        const dd linv = dd_recip(dd_sqrt(d));
        for (int i = j + 1; i < m; ++i) at(i, j) = dd_mul(at(i, j), linv);
        for (int i = j + 1; i < m; ++i)
            for (int c = j + 1; c <= i && c < n; ++c) at(i, c) = dd_fnma(at(i, j), at(c, j), at(i, c));
    }
    dd sab{0, 0}, sbb{0, 0};
    for (int i = 0; i < n; ++i) {
        sab = dd_add(sab, dd_mul(at(n, i), at(n + 1, i)));
        sbb = dd_add(sbb, dd_sqr(at(n + 1, i)));
    }
    const dd mu = dd_div(sab, sbb);
    dd srr{0, 0};
    for (int i = 0; i < n; ++i) {
        const dd r = dd_sub(at(n, i), dd_mul(mu, at(n + 1, i)));
        srr = dd_add(srr, dd_sqr(r));
    }
    const dd s2 = dd_mul_d(srr, 1.0 / n);  // n is not a power of two in general: use a true division
    const dd s2d = dd_div(srr, dd{(double)n, 0.0});
    (void)s2;
    *minpiv = mp;
    return 0.5 * n * (log(s2d.hi) + s2d.lo / s2d.hi) + 0.5 * logdet;
}
void dd_nll_batch(const double* X, const double* y, int n, int k, const double* theta, const double* p, long C,
                  double* nll, double* minpiv) {
#pragma omp parallel for schedule(dynamic, 1)
    for (long c = 0; c < C; ++c) nll[c] = dd_nll_one(X, y, n, k, theta + c * k, p + c * k, minpiv + c);
}
}
