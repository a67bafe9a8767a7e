/* Extended-precision adjudicator for the Kriging likelihood -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference (mylibs/kriging.py:66-116) evaluates every candidate in fp64 through an LU inverse; for
 * cond(R) >~ 1e12 its value is rounding noise.  This file restates the same quantity
 *     R_ij = exp(-sum_d theta_d |x_id - x_jd|^p_d)                      (kriging.py:70-77)
 *     mu = 1'R^-1 y / 1'R^-1 1,  sigma^2 = (y-mu)'R^-1(y-mu)/n          (kriging.py:86-93,106)
 *     nll = n/2 ln sigma^2 + 1/2 ln det R                                (kriging.py:111)
 * twice:
 *   quad_nll_batch  -- IEEE binary128 (__float128, libquadmath) from the fp64 inputs: entries, Cholesky,
 *                      solves and logarithms all carry 113 bits, so the result is the exact value to
 *                      ~1e-34 * cond(R).  This is the 'truth' of App. E rule (iii), ~100x faster than mpmath
 *                      (checked against oracle/mp_oracle.py in tests/test_oracle_golden.py).
 *   f64_nll_batch   -- plain fp64 left-looking Cholesky of the same formula (libm exp/pow); also returns
 *                      the smallest pivot, used to map which candidates of a grid are ill-conditioned.
 * Both are OpenMP-parallel over candidates.  Nothing under surrogate-models_b200/ links or loads this.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC oracle/quad_oracle.c -o oracle/_build/libquad_oracle.so -lquadmath -lm
 */
#include <math.h>
#include <quadmath.h>
#include <stdlib.h>
#include <string.h>

typedef __float128 q128;

/* theta is passed as fp64 (the reference forms 10.**e in fp64, kriging.py:133-136) */
static double quad_nll_one(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                           double* min_pivot, double* mu_out, double* s2_out, q128* L, q128* a, q128* b) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            q128 s = 0;
            for (int d = 0; d < k; ++d) {
                q128 dx = fabsq((q128)X[i * k + d] - (q128)X[j * k + d]);
                q128 pw = (dx == 0) ? (q128)0 : powq(dx, (q128)p[d]);
                s += (q128)theta[d] * pw;
            }
            L[i * n + j] = expq(-s);
        }
    q128 minp = 1e300Q, logdet = 0;
    for (int j = 0; j < n; ++j) {
        q128 d = L[j * n + j];
        for (int t = 0; t < j; ++t) d -= L[j * n + t] * L[j * n + t];
        if (d < minp) minp = d;
        if (!(d > 0)) {
            *min_pivot = (double)d;
            return INFINITY;
        }
        q128 ljj = sqrtq(d);
        logdet += logq(d);
        L[j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            q128 s = L[i * n + j];
            for (int t = 0; t < j; ++t) s -= L[i * n + t] * L[j * n + t];
            L[i * n + j] = s / ljj;
        }
    }
    for (int i = 0; i < n; ++i) {
        q128 sa = y[i], sb = 1;
        for (int t = 0; t < i; ++t) {
            sa -= L[i * n + t] * a[t];
            sb -= L[i * n + t] * b[t];
        }
        a[i] = sa / L[i * n + i];
        b[i] = sb / L[i * n + i];
    }
    q128 sab = 0, sbb = 0;
    for (int i = 0; i < n; ++i) {
        sab += a[i] * b[i];
        sbb += b[i] * b[i];
    }
    q128 mu = sab / sbb, srr = 0;
    for (int i = 0; i < n; ++i) {
        q128 r = a[i] - mu * b[i];
        srr += r * r;
    }
    q128 s2 = srr / n;
    *min_pivot = (double)minp;
    if (mu_out) *mu_out = (double)mu;
    if (s2_out) *s2_out = (double)s2;
    return (double)(0.5Q * n * logq(s2) + 0.5Q * logdet);
}

int quad_nll_batch(const double* X, const double* y, int n, int k, const double* theta, const double* p, long C,
                   double* nll, double* min_pivot, double* mu, double* s2) {
#pragma omp parallel
    {
        q128* L = (q128*)malloc(sizeof(q128) * ((size_t)n * n + 2 * n));
        q128 *a = L + (size_t)n * n, *b = a + n;
#pragma omp for schedule(dynamic, 1)
        for (long c = 0; c < C; ++c) {
            double mp = 0, m = NAN, s = NAN;
            nll[c] = quad_nll_one(X, y, n, k, theta + c * k, p + c * k, &mp, &m, &s, L, a, b);
            if (min_pivot) min_pivot[c] = mp;
            if (mu) mu[c] = m;
            if (s2) s2[c] = s;
        }
        free(L);
    }
    return 0;
}

static double f64_nll_one(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                          double* min_pivot, double* L, double* a, double* b) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0;
            for (int d = 0; d < k; ++d) s += theta[d] * pow(fabs(X[i * k + d] - X[j * k + d]), p[d]);
            L[i * n + j] = exp(-s);
        }
    double minp = 1e300, logdet = 0;
    for (int j = 0; j < n; ++j) {
        double d = L[j * n + j];
        for (int t = 0; t < j; ++t) d -= L[j * n + t] * L[j * n + t];
        if (d < minp) minp = d;
        if (!(d > 0)) {
            *min_pivot = d;
            return INFINITY;
        }
        double ljj = sqrt(d);
        logdet += log(d);
        L[j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            double s = L[i * n + j];
            for (int t = 0; t < j; ++t) s -= L[i * n + t] * L[j * n + t];
            L[i * n + j] = s / ljj;
        }
    }
    for (int i = 0; i < n; ++i) {
        double sa = y[i], sb = 1;
        for (int t = 0; t < i; ++t) {
            sa -= L[i * n + t] * a[t];
            sb -= L[i * n + t] * b[t];
        }
        a[i] = sa / L[i * n + i];
        b[i] = sb / L[i * n + i];
    }
    double sab = 0, sbb = 0;
    for (int i = 0; i < n; ++i) {
        sab += a[i] * b[i];
        sbb += b[i] * b[i];
    }
    double mu = sab / sbb, srr = 0;
    for (int i = 0; i < n; ++i) {
        double r = a[i] - mu * b[i];
        srr += r * r;
    }
    *min_pivot = minp;
    return 0.5 * n * log(srr / n) + 0.5 * logdet;
}

int f64_nll_batch(const double* X, const double* y, int n, int k, const double* theta, const double* p, long C,
                  double* nll, double* min_pivot) {
#pragma omp parallel
    {
        double* L = (double*)malloc(sizeof(double) * ((size_t)n * n + 2 * n));
        double *a = L + (size_t)n * n, *b = a + n;
#pragma omp for schedule(static, 256)
        for (long c = 0; c < C; ++c) {
            double mp = 0;
            nll[c] = f64_nll_one(X, y, n, k, theta + c * k, p + c * k, &mp, L, a, b);
            if (min_pivot) min_pivot[c] = mp;
        }
        free(L);
    }
    return 0;
}

/* Exact likelihood of the fp64-ROUNDED correlation matrix: entries formed exactly as the reference forms them
 * (fp64 pow / multiply / add / exp through libm, kriging.py:70-77), everything after that in binary128.
 * This is the function every fp64 implementation of the path (the reference's LU, LAPACK potrf, the GPU's
 * Cholesky) approximates; it separates 'the matrix the reference holds is indefinite' from 'the
 * factorisation lost the pivot to rounding'. */
static double quad_nll_rounded_one(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                                   double* min_pivot, q128* L, q128* a, q128* b) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0;
            for (int d = 0; d < k; ++d) s += theta[d] * pow(fabs(X[i * k + d] - X[j * k + d]), p[d]);
            L[i * n + j] = (q128)exp(-s);
        }
    q128 minp = 1e300Q, logdet = 0;
    for (int j = 0; j < n; ++j) {
        q128 d = L[j * n + j];
        for (int t = 0; t < j; ++t) d -= L[j * n + t] * L[j * n + t];
        if (d < minp) minp = d;
        if (!(d > 0)) {
            *min_pivot = (double)d;
            return INFINITY;
        }
        q128 ljj = sqrtq(d);
        logdet += logq(d);
        L[j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            q128 s = L[i * n + j];
            for (int t = 0; t < j; ++t) s -= L[i * n + t] * L[j * n + t];
            L[i * n + j] = s / ljj;
        }
    }
    for (int i = 0; i < n; ++i) {
        q128 sa = y[i], sb = 1;
        for (int t = 0; t < i; ++t) {
            sa -= L[i * n + t] * a[t];
            sb -= L[i * n + t] * b[t];
        }
        a[i] = sa / L[i * n + i];
        b[i] = sb / L[i * n + i];
    }
    q128 sab = 0, sbb = 0;
    for (int i = 0; i < n; ++i) {
        sab += a[i] * b[i];
        sbb += b[i] * b[i];
    }
    q128 mu = sab / sbb, srr = 0;
    for (int i = 0; i < n; ++i) {
        q128 r = a[i] - mu * b[i];
        srr += r * r;
    }
    *min_pivot = (double)minp;
    return (double)(0.5Q * n * logq(srr / n) + 0.5Q * logdet);
}

int quad_nll_rounded_batch(const double* X, const double* y, int n, int k, const double* theta, const double* p,
                           long C, double* nll, double* min_pivot) {
#pragma omp parallel
    {
        q128* L = (q128*)malloc(sizeof(q128) * ((size_t)n * n + 2 * n));
        q128 *a = L + (size_t)n * n, *b = a + n;
#pragma omp for schedule(dynamic, 1)
        for (long c = 0; c < C; ++c) {
            double mp = 0;
            nll[c] = quad_nll_rounded_one(X, y, n, k, theta + c * k, p + c * k, &mp, L, a, b);
            if (min_pivot) min_pivot[c] = mp;
        }
        free(L);
    }
    return 0;
}
