"""Quick GPU sanity run of every host-buffer C-ABI entry point against the numpy oracle (dev tool)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
from sm_b200 import _cabi as C  # noqa: E402
from oracle import numpy_oracle as orc  # noqa: E402

L = C.lib()
rng = np.random.default_rng(0)


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    fin = np.isfinite(a) & np.isfinite(b)
    same_nonfinite = np.array_equal(np.isfinite(a), np.isfinite(b))
    e = np.max(np.abs(a[fin] - b[fin]) / np.maximum(1, np.abs(b[fin]))) if fin.any() else 0.0
    return e, same_nonfinite


def check_batch(n, k, C_=24):
    X = rng.uniform(0, 10, (n, k))
    y = np.sin(X.sum(axis=1)) + 0.1 * X[:, 0] ** 2
    e = rng.uniform(-2, 0.5, (C_, k))
    p = rng.uniform(1, 2, (C_, k))
    p[0] = 2.0
    p[1] = 1.0
    th = C.f64(10. ** e)
    nll = np.empty(C_)
    bv = np.empty(1)
    bi = np.empty(1, dtype=np.int64)
    t0 = time.time()
    C.check(L.sm_krig_nll_batch_host(C.hptr(C.f64(X)), C.hptr(C.f64(y)), n, k, C.hptr(th), C.hptr(C.f64(p)), C_, 0,
                                     C.hptr(nll), C.hptr(bv), C.hptr(bi)))
    dt = time.time() - t0
    ref = np.array([orc.nll_chol(X, y, th[c], p[c]) for c in range(C_)])
    err, nf = relerr(nll, ref)
    print(f'nll_batch n={n:5d} k={k} C={C_}: relerr {err:.2e} nonfinite_match {nf} argmin {bi[0]} vs {orc.first_min_vec(ref)} ({dt*1e3:.1f} ms)', flush=True)
    return err


def check_grid(n, k, T, P):
    X = rng.uniform(0, 10, (n, k))
    y = np.sin(X.sum(axis=1)) + 0.1 * X[:, 0] ** 2
    th_axis, p_axis = orc.grid_axes(T, P)
    th_lin = np.array([10. ** float(e) for e in th_axis])
    Cn = (T ** k) * (P ** k)
    nll = np.empty(Cn)
    bv = np.empty(1)
    bi = np.empty(1, dtype=np.int64)
    t0 = time.time()
    C.check(L.sm_krig_nll_grid_host(C.hptr(C.f64(X)), C.hptr(C.f64(y)), n, k, C.hptr(th_lin), T, C.hptr(C.f64(p_axis)), P,
                                    0, Cn, C.hptr(nll), C.hptr(bv), C.hptr(bi)))
    dt = time.time() - t0
    e, q = orc.grid_candidates(k, T, P)
    sel = np.unique(np.concatenate([np.arange(0, Cn, max(1, Cn // 200)), [bi[0]]]))
    ref = np.array([orc.nll_chol(X, y, 10. ** e[c], q[c]) for c in sel])
    conds = np.array([orc.cond2(orc.corr_matrix(X, 10. ** e[c], q[c])) for c in sel])
    well = conds < 1e8
    err, _ = relerr(nll[sel][well], ref[well])
    nf_gpu, nf_ref = np.sum(~np.isfinite(nll[sel])), np.sum(~np.isfinite(ref))
    print(f'nll_grid  n={n:5d} k={k} C={Cn}: relerr(cond<1e8, {well.sum()} pts) {err:.2e}; inf gpu/ref {nf_gpu}/{nf_ref}; '
          f'argmin {bi[0]} val {bv[0]:.10g} first_min(gpu vec) {orc.first_min_vec(nll)} ({dt*1e3:.1f} ms, {Cn/dt:.3g} evals/s e2e)', flush=True)


def check_predict(n, k, Nq=1000, mse=True):
    X = rng.uniform(0, 10, (n, k))
    y = np.sin(X.sum(axis=1)) + 0.1 * X[:, 0] ** 2
    th = np.full(k, 0.3)
    p = np.full(k, 1.7)
    Xq = rng.uniform(-1, 11, (Nq, k))
    Xq[:4] = X[:4]
    yo = np.empty(Nq)
    mo = np.empty(Nq) if mse else None
    sc = np.empty(4)
    t0 = time.time()
    C.check(L.sm_krig_fit_predict_host(C.hptr(C.f64(X)), C.hptr(C.f64(y)), n, k, C.hptr(th), C.hptr(p), C.hptr(C.f64(Xq)), Nq,
                                       C.hptr(yo), C.hptr(mo), C.hptr(sc)))
    dt = time.time() - t0
    f = orc.fit_chol(X, y, th, p)
    if mse:
        yr, mr = orc.predict_chol(X, y, th, p, Xq, True)
        em = np.max(np.abs(mo - mr)) / f['sigma2']
    else:
        yr = orc.predict_chol(X, y, th, p, Xq)
        em = 0.0
    ey = np.max(np.abs(yo - yr) / np.maximum(1, np.abs(yr)))
    nl = orc.nll_chol(X, y, th, p)
    print(f'predict   n={n:5d} k={k} Nq={Nq} mse={mse}: y relerr {ey:.2e} mse abs/sigma2 {em:.2e}; nll {sc[0]:.12g} vs {nl:.12g}; '
          f'mu {sc[1]:.12g} vs {f["mu"]:.12g} info {sc[3]} ({dt*1e3:.1f} ms)', flush=True)


def check_rbf(n, k, kind, a, Nq=500):
    X = rng.uniform(0, 10, (n, k))
    y = np.sin(X[:, 0]) + 0.05 * X[:, -1] ** 2
    Xq = rng.uniform(0, 10, (Nq, k))
    coeff = np.empty(n)
    info = np.zeros(1, dtype=np.int32)
    t0 = time.time()
    C.check(L.sm_rbf_fit_host(C.hptr(C.f64(X)), C.hptr(C.f64(y)), n, k, kind, a, C.hptr(coeff), C.hptr(info)))
    dt = time.time() - t0
    cr = orc.rbf_fit(X, y, kind, a)
    yo = np.empty(Nq)
    C.check(L.sm_rbf_predict_host(C.hptr(C.f64(X)), n, k, kind, a, C.hptr(coeff), C.hptr(C.f64(Xq)), Nq, C.hptr(yo)))
    yr = orc.rbf_predict(X, cr, kind, a, Xq)
    ec = np.max(np.abs(coeff - cr)) / np.max(np.abs(cr))
    ey = np.max(np.abs(yo - yr) / np.maximum(1, np.abs(yr)))
    print(f'rbf       n={n:5d} k={k} kind={kind}: coeff relerr {ec:.2e} pred relerr {ey:.2e} info {info[0]} cond {orc.cond2(orc.rbf_matrix(X, kind, a)):.2e} (fit {dt*1e3:.1f} ms)', flush=True)


if __name__ == '__main__':
    print('version', L.sm_b200_version())
    for n, k in [(8, 1), (13, 2), (36, 2), (64, 2), (72, 3), (100, 2), (200, 4), (256, 2), (300, 8)]:
        check_batch(n, k)
    check_grid(8, 1, 8, 5)
    check_grid(36, 2, 8, 5)
    check_grid(64, 2, 8, 5)
    check_grid(64, 2, 40, 25)
    for n, k, m in [(8, 1, True), (36, 2, True), (36, 2, False), (64, 3, True), (100, 2, True), (150, 5, True), (150, 5, False)]:
        check_predict(n, k, mse=m)
    for kind, a in [(2, 1.0), (0, 0.17), (1, 1.0), (3, 0.5), (4, 1.3)]:
        check_rbf(36, 2, kind, a)
    check_rbf(500, 3, 2, 1.0)
    check_rbf(300, 3, 3, 0.5)
    print('launches', C.launch_count())
