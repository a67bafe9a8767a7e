"""CPU oracle for the Kriging / RBF hot path -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference algorithm (juri117/surrogate-models,
``mylibs/kriging.py``, ``mylibs/likeli_optimizer.py``, ``mylibs/rbf.py``).  It exists to check
the CUDA path; nothing under ``surrogate-models_b200/`` may import it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
use it, and there only as the checker / the CPU arm being timed.

Parity status: the reference ships no tests, so the oracle is pinned against
outputs of the *reference itself*, executed in the build container by ``oracle/gen_golden.py``
and committed under ``tests/golden/`` (see ``tests/test_oracle_golden.py``).  The one in-repo
known answer (``examples/kriging_r2.py:77-79``) is part of those fixtures.

Two formulations are provided for the likelihood:

* ``*_lu``   -- the reference's own algorithm: ``np.linalg.inv`` + ``np.linalg.slogdet(...)[1]``
               (``mylibs/kriging.py:79,86-93,95-116``).  This is what ``cpu_baseline`` times.
* ``*_chol`` -- the Cholesky formulation the GPU kernels compute (forward solves of y and 1).
               Agrees with ``*_lu`` to <= 1e-12 for cond(R) <~ 1e8.

``*_loops`` variants are pure-Python scalar loops that mirror the reference statement by
statement (``math.exp``, scalar ``**``) and are bit-identical to it; small cases only.
"""
import math

import numpy as np

# ----------------------------------------------------------------------------------------------
# correlation matrix  (mylibs/kriging.py:66-77)
# ----------------------------------------------------------------------------------------------


def corr_matrix(X, theta, p):
    """R_ij = exp(-sum_d theta_d |x_id - x_jd|**p_d)   (mylibs/kriging.py:70-77), vectorised."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    theta = np.asarray(theta, dtype=np.float64).ravel()
    p = np.asarray(p, dtype=np.float64).ravel()
    n, k = X.shape
    s = np.zeros((n, n))
    for d in range(k):  # same left-to-right accumulation order as the reference's `sum +=`
        s = s + theta[d] * np.abs(X[:, d][:, None] - X[:, d][None, :]) ** p[d]
    return np.exp(-s)


def corr_matrix_loops(X, theta, p):
    """Statement-by-statement port of mylibs/kriging.py:70-77 (scalar pow, math.exp)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    n, k = X.shape
    R = np.zeros((n, n))
    for i in range(n):
        for j in range(i, n):
            s = 0.
            for d in range(k):
                s += theta[d] * (abs(X[i][d] - X[j][d]) ** p[d])
            R[i][j] = math.exp(-s)
            R[j][i] = R[i][j]
    return R


# ----------------------------------------------------------------------------------------------
# concentrated negative log-likelihood
# ----------------------------------------------------------------------------------------------


def nll_from_R_lu(R, y):
    """Reference algorithm on a given R: inv (kriging.py:79), mu (:86-93), slogdet[1] (:100),
    sigma^2 (:106), nll (:111) with the inf guards of :101-110.  Returns (nll, mu, sigma2)."""
    n = R.shape[0]
    y = np.asarray(y, dtype=np.float64)
    Rinv = np.linalg.inv(R)  # a singular matrix raises; the reference keeps stale state there
    one = np.ones(n)
    mu = (one @ Rinv @ y) / (one @ Rinv @ one)
    ln_det = np.linalg.slogdet(R)[1]
    if np.isnan(ln_det):
        return float('inf'), mu, float('nan')
    r = y - one * mu
    sigma2 = (r @ Rinv @ r) / n
    if sigma2 < 0.:
        return float('inf'), mu, sigma2
    with np.errstate(divide='ignore'):
        nll = (-1) * (-(n / 2) * np.log(sigma2) - 0.5 * ln_det)
    return float(nll), float(mu), float(sigma2)


def nll_from_R_chol(R, y):
    """Cholesky formulation (what the CUDA kernels compute): R = L L^T, a = L^-1 y, b = L^-1 1,
    mu = a.b / b.b, r = a - mu b, sigma^2 = r.r / n, nll = n/2 ln sigma^2 + sum ln L_ii.
    Not numerically SPD  ->  +inf (policy documented in DESIGN.md).  Returns (nll, mu, sigma2)."""
    import scipy.linalg as sla
    n = R.shape[0]
    y = np.asarray(y, dtype=np.float64)
    try:
        L = np.linalg.cholesky(R)
    except np.linalg.LinAlgError:
        return float('inf'), float('nan'), float('nan')
    ab = sla.solve_triangular(L, np.stack([y, np.ones(n)], axis=1), lower=True)
    a, b = ab[:, 0], ab[:, 1]
    mu = (a @ b) / (b @ b)
    r = a - mu * b
    sigma2 = (r @ r) / n
    with np.errstate(divide='ignore'):
        nll = (n / 2) * np.log(sigma2) + np.sum(np.log(np.diag(L)))
    if np.isnan(nll):
        return float('inf'), float(mu), float(sigma2)
    return float(nll), float(mu), float(sigma2)


def nll_lu(X, y, theta, p):
    """update_param + calc_likelihood of the reference (kriging.py:53-63,95-116)."""
    return nll_from_R_lu(corr_matrix(X, theta, p), y)[0]


def nll_chol(X, y, theta, p):
    return nll_from_R_chol(corr_matrix(X, theta, p), y)[0]


def nll_lu_loops(X, y, theta, p):
    """Bit-faithful (scalar-loop) restatement of update_param + calc_likelihood."""
    return nll_from_R_lu(corr_matrix_loops(X, theta, p), y)[0]


def nll_opti_exp(X, y, params, form='lu'):
    """_calc_likelihood_opti_exp (kriging.py:130-141): params = [log10 theta_1..k, p_1..k];
    NaN in params -> NaN."""
    params = np.asarray(params, dtype=np.float64)
    if np.isnan(params).any():
        return float('nan')
    k = params.shape[0] // 2
    theta = [10. ** e for e in params[:k]]  # scalar pow like the reference (:133-136)
    f = nll_lu if form == 'lu' else nll_chol
    try:
        return f(X, y, theta, params[k:])
    except np.linalg.LinAlgError:
        # exactly singular LU: the reference swallows the exception and reports the likelihood of the
        # PREVIOUS candidate (stale _cor_mat/_core_mat_inv, kriging.py:78-83).  The oracle -- like the
        # GPU path -- reports +inf instead (DESIGN.md "Numerical contract").
        return float('inf')


def nll_batch(X, y, log10_theta, p, form='lu'):
    """Likelihood of every row of a candidate table (C,k)+(C,k).  Singular LU -> +inf."""
    log10_theta = np.atleast_2d(np.asarray(log10_theta, dtype=np.float64))
    p = np.atleast_2d(np.asarray(p, dtype=np.float64))
    out = np.empty(log10_theta.shape[0])
    for c in range(out.shape[0]):
        out[c] = nll_opti_exp(X, y, np.concatenate([log10_theta[c], p[c]]), form)
    return out


# ----------------------------------------------------------------------------------------------
# candidate enumeration  (mylibs/likeli_optimizer.py:50-97)
# ----------------------------------------------------------------------------------------------


def grid_axes(theta_sections, p_sections):
    """likeli_optimizer.py:53-55: p axis linspace(1,2,P); theta axis = EXPONENTS linspace(-5,5,T)."""
    return np.linspace(-5, 5, num=theta_sections), np.linspace(1., 2., num=p_sections)


def grid_candidates(k, theta_sections, p_sections, lo=0, hi=None):
    """Rows [lo, hi) of the candidate table in the reference's enumeration order: outer p-odometer,
    inner theta-odometer, dimension 0 the fastest digit (increase_i, likeli_optimizer.py:89-97).
    Returns (log10_theta[C,k], p[C,k])."""
    th_axis, p_axis = grid_axes(theta_sections, p_sections)
    T, P = theta_sections, p_sections
    total = (T ** k) * (P ** k)
    hi = total if hi is None else hi
    c = np.arange(lo, hi, dtype=np.int64)
    ip, it = np.divmod(c, T ** k)
    e = np.empty((c.shape[0], k))
    q = np.empty((c.shape[0], k))
    for d in range(k):
        e[:, d] = th_axis[(it // (T ** d)) % T]
        q[:, d] = p_axis[(ip // (P ** d)) % P]
    return e, q


def first_min(values):
    """Strict '<' scan keeping the FIRST minimum; NaN never wins; all inf/NaN -> -1
    (likeli_optimizer.py:51,79-81)."""
    best, idx = float('inf'), -1
    for i, v in enumerate(values):
        if v < best:
            best, idx = v, i
    return idx


def first_min_vec(values):
    v = np.asarray(values, dtype=np.float64)
    w = np.where(np.isnan(v), np.inf, v)
    i = int(np.argmin(w))  # np.argmin returns the first occurrence
    return i if w[i] < np.inf else -1


# ----------------------------------------------------------------------------------------------
# fit + predict  (mylibs/kriging.py:53-63, 206-220)
# ----------------------------------------------------------------------------------------------


def psi_vectors(X, theta, p, Xq):
    """psi[q,i] = exp(-sum_d theta_d |x_id - xq_d|**p_d)  (kriging.py:214-218)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    Xq = np.asarray(Xq, dtype=np.float64).reshape(-1, X.shape[1])
    s = np.zeros((Xq.shape[0], X.shape[0]))
    for d in range(X.shape[1]):
        s = s + theta[d] * np.abs(X[:, d][None, :] - Xq[:, d][:, None]) ** p[d]
    return np.exp(-s)


def predict_lu(X, y, theta, p, Xq):
    """Reference predict (kriging.py:206-220): mu + (psi^T R^-1)(y - 1 mu), R^-1 = np.linalg.inv."""
    y = np.asarray(y, dtype=np.float64)
    R = corr_matrix(X, theta, p)
    Rinv = np.linalg.inv(R)
    one = np.ones(R.shape[0])
    mu = (one @ Rinv @ y) / (one @ Rinv @ one)
    psi = psi_vectors(X, theta, p, Xq)
    return mu + (psi @ Rinv) @ (y - one * mu)


def predict_lu_loops(X, y, theta, p, Rinv, mu, x_pred):
    """Statement-by-statement port of Kriging.predict (kriging.py:206-220) on the state update_param left behind
    (Rinv = _core_mat_inv, mu = _mu): interpreted psi loop with scalar pow / math.exp, then the O(n^2) product
    psi^T R^-1 (y - 1 mu) -- per query, as the reference does.  This is what bench.py's prediction cpu_baseline times."""
    n, k = X.shape
    one = np.ones((n, 1)).flatten()
    psi = np.ones((n, 1)).flatten()
    for i in range(0, n):
        s = 0.
        for ik in range(0, k):
            s += theta[ik] * (abs(X[i][ik] - x_pred[ik]) ** p[ik])
        psi[i] = math.exp(-s)
    return mu + np.transpose(psi) @ Rinv @ (y - one * mu)


def rbf_predict_loops(X, coeff, a, x_pred):
    """Statement-by-statement port of RBF.predict with the gaussian kernel (rbf.py:85-103,111-112)."""
    res = 0.
    for i in range(0, X.shape[0]):
        rad_sum = 0.
        for ik in range(0, X.shape[1]):
            rad_sum += (x_pred[ik] - X[i][ik]) ** 2.
        radius = math.sqrt(rad_sum)
        res += coeff[i] * math.e ** (-((a * radius) ** 2))
    return res


def fit_chol(X, y, theta, p):
    """Fitted state the GPU path keeps: L, M = L^-1, a = L^-1 y, b = L^-1 1, mu, sigma2,
    w = R^-1 (y - 1 mu) = M^T (a - mu b)."""
    import scipy.linalg as sla
    y = np.asarray(y, dtype=np.float64)
    R = corr_matrix(X, theta, p)
    n = R.shape[0]
    L = np.linalg.cholesky(R)
    M = sla.solve_triangular(L, np.eye(n), lower=True)
    a, b = M @ y, M @ np.ones(n)
    mu = (a @ b) / (b @ b)
    r = a - mu * b
    return dict(R=R, L=L, M=M, a=a, b=b, mu=mu, sigma2=(r @ r) / n, w=M.T @ r)


def predict_chol(X, y, theta, p, Xq, return_mse=False):
    """Mean (and ordinary-Kriging MSE) in the Cholesky form.  MSE (not in the reference --
    'parity unpinned', pinned against oracle/mp_oracle.py instead):
        s^2(x) = sigma^2 [ 1 - c.c + (1 - b.c)^2 / (b.b) ],  c = L^-1 psi,  clamped at 0."""
    f = fit_chol(X, y, theta, p)
    psi = psi_vectors(X, theta, p, Xq)
    yhat = f['mu'] + psi @ f['w']
    if not return_mse:
        return yhat
    c = psi @ f['M'].T
    u = 1. - c @ f['b']
    mse = f['sigma2'] * (1. - np.sum(c * c, axis=1) + u * u / (f['b'] @ f['b']))
    return yhat, np.maximum(mse, 0.)


# ----------------------------------------------------------------------------------------------
# RBF  (mylibs/rbf.py:67-118)
# ----------------------------------------------------------------------------------------------

RBF_KINDS = {'lin': 0, 'cubic': 1, 'gaus': 2, 'multi-quadratic': 3, 'mq': 3,
             'inverse-multi-quadratic': 4, 'imq': 4}


def rbf_phi(kind, a, r):
    """The five kernels of rbf.py:105-118, incl. the reference's own `imq` = (1+r^2)^(a/2)."""
    r = np.asarray(r, dtype=np.float64)
    if kind == 0:
        return r
    if kind == 1:
        return r ** 3
    if kind == 2:
        return math.e ** (-((a * r) ** 2))
    if kind == 3:
        return np.sqrt(1 + (a * r) ** 2)
    if kind == 4:
        return (1 + r ** 2) ** (a / 2)
    raise ValueError(kind)


def rbf_matrix(X, kind, a):
    """Phi_ij = phi(a, ||x_i - x_j||_2)  (rbf.py:70-78)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    s = np.zeros((X.shape[0], X.shape[0]))
    for d in range(X.shape[1]):
        s = s + (X[:, d][:, None] - X[:, d][None, :]) ** 2.
    return rbf_phi(kind, a, np.sqrt(s))


def rbf_fit(X, y, kind, a):
    """coeff = np.linalg.solve(Phi, y)  (rbf.py:79); singular Phi raises LinAlgError."""
    return np.linalg.solve(rbf_matrix(X, kind, a), np.asarray(y, dtype=np.float64))


def rbf_predict(X, coeff, kind, a, Xq):
    """sum_i coeff_i phi(a, ||x - x_i||)  (rbf.py:85-103)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    Xq = np.asarray(Xq, dtype=np.float64).reshape(-1, X.shape[1])
    s = np.zeros((Xq.shape[0], X.shape[0]))
    for d in range(X.shape[1]):
        s = s + (Xq[:, d][:, None] - X[:, d][None, :]) ** 2.
    return rbf_phi(kind, a, np.sqrt(s)) @ np.asarray(coeff, dtype=np.float64)


def cond2(R):
    return float(np.linalg.cond(R))
