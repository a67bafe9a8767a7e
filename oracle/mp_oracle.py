"""High-precision (mpmath, 60 digits) adjudicator -- TEST INFRASTRUCTURE ONLY.

Used for (i) the ordinary-Kriging MSE, which the reference does not implement (Kriging.predict returns
the mean only, mylibs/kriging.py:206-220) -> 'parity unpinned' for that quantity, so its fixture comes from
exact arithmetic on the textbook formula, and (ii) ill-conditioned candidates where the reference's
LU-inverse path itself is inaccurate.

    python oracle/mp_oracle.py      # writes tests/golden/mp_kriging_r3.npz
"""
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60


def corr(X, theta, p):
    n, k = X.shape
    R = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            s = mp.mpf(0)
            for d in range(k):
                s += mp.mpf(float(theta[d])) * abs(mp.mpf(float(X[i, d])) - mp.mpf(float(X[j, d]))) ** mp.mpf(float(p[d]))
            R[i, j] = mp.e ** (-s)
    return R


def fit(X, y, theta, p):
    n = X.shape[0]
    R = corr(X, theta, p)
    Rinv = R ** -1
    one = mp.matrix([1] * n)
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    mu = (one.T * Rinv * yv)[0] / (one.T * Rinv * one)[0]
    r = yv - one * mu
    s2 = (r.T * Rinv * r)[0] / n
    nll = mp.mpf(n) / 2 * mp.log(s2) + mp.log(mp.det(R)) / 2
    return dict(R=R, Rinv=Rinv, mu=mu, s2=s2, nll=nll, w=Rinv * r, one=one)


def nll_chol(X, y, theta, p, dps=100):
    """Likelihood by a Cholesky factorisation in `dps`-digit arithmetic (kriging.py:66-116 in Cholesky form): the
    adjudicator for candidates beyond binary128 (cond > 1e30).  Returns a float (+inf only if a pivot is <= 0 even
    at this precision, which does not happen for distinct samples)."""
    old = mp.mp.dps
    mp.mp.dps = dps
    try:
        n = X.shape[0]
        R = corr(X, theta, p)
        L = [[mp.mpf(0)] * n for _ in range(n)]
        logdet = mp.mpf(0)
        for j in range(n):
            d = R[j, j] - sum(L[j][t] * L[j][t] for t in range(j))
            if d <= 0:
                return float('inf')
            ljj = mp.sqrt(d)
            logdet += mp.log(d)
            L[j][j] = ljj
            for i in range(j + 1, n):
                L[i][j] = (R[i, j] - sum(L[i][t] * L[j][t] for t in range(j))) / ljj
        a, b = [], []
        for i in range(n):
            a.append((mp.mpf(float(y[i])) - sum(L[i][t] * a[t] for t in range(i))) / L[i][i])
            b.append((1 - sum(L[i][t] * b[t] for t in range(i))) / L[i][i])
        mu = sum(u * v for u, v in zip(a, b)) / sum(v * v for v in b)
        s2 = sum((u - mu * v) ** 2 for u, v in zip(a, b)) / n
        return float(mp.mpf(n) / 2 * mp.log(s2) + logdet / 2)
    finally:
        mp.mp.dps = old


def predict(X, f, theta, p, Xq):
    n, k = X.shape
    ys, ms = [], []
    den = (f['one'].T * f['Rinv'] * f['one'])[0]
    for q in range(Xq.shape[0]):
        psi = mp.matrix(n, 1)
        for i in range(n):
            s = mp.mpf(0)
            for d in range(k):
                s += mp.mpf(float(theta[d])) * abs(mp.mpf(float(X[i, d])) - mp.mpf(float(Xq[q, d]))) ** mp.mpf(float(p[d]))
            psi[i] = mp.e ** (-s)
        ys.append(f['mu'] + (psi.T * f['w'])[0])
        t = f['Rinv'] * psi
        u = 1 - (f['one'].T * t)[0]
        ms.append(f['s2'] * (1 - (psi.T * t)[0] + u * u / den))
    return np.array([float(v) for v in ys]), np.array([float(v) for v in ms])


if __name__ == '__main__':
    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, '..', 'tests', 'golden', 'c5_kriging_r3.npz'))
    X, y = g['X'], g['y']
    xq = np.concatenate([g['xq'][:40], g['xq'][-11:]])  # random points, 8 sample points, 3 named points
    out = {}
    for tag, theta, p in (('fixed', [0.05, 0.05], [1.9, 1.9]), ('rough', [0.3, 0.8], [1.3, 1.7])):
        f = fit(X, y, theta, p)
        yy, mm = predict(X, f, theta, p, xq)
        out.update({f'theta_{tag}': np.array(theta), f'p_{tag}': np.array(p), f'y_{tag}': yy, f'mse_{tag}': mm,
                    f'mu_{tag}': float(f['mu']), f's2_{tag}': float(f['s2']), f'nll_{tag}': float(f['nll'])})
        print(tag, float(f['mu']), float(f['s2']), float(f['nll']), mm[-3:])
    np.savez_compressed(os.path.join(here, '..', 'tests', 'golden', 'mp_kriging_r3.npz'), X=X, y=y, xq=xq,
                        dps=mp.mp.dps, **out)
