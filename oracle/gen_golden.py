"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (read-only /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py            # everything except the full 1e6-candidate C2 grids
    python oracle/gen_golden.py --full-c2  # + both full C2 grids (about 15 min each on 8 cores)
    python oracle/gen_golden.py --only illcond  # ill-conditioned corner of the C2 grids (binary128 / mpmath adjudication)

Test infrastructure; the product never imports this.  Inputs are produced with the reference's own
generators (mylibs/halton.py, mylibs/structured_sample.py, myutils/samples.py) and stored beside
the outputs, so the fixtures are self-contained.
"""
import argparse
import contextlib
import io
import math
import os
import sys
import time

import numpy as np

REF = os.environ.get('SM_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

from mylibs.kriging import Kriging  # noqa: E402  (reference)
from mylibs.rbf import RBF  # noqa: E402
from mylibs.likeli_optimizer import LikeliOptimizer  # noqa: E402
from mylibs.halton import Halton  # noqa: E402
from mylibs.structured_sample import StructuredSample  # noqa: E402
from myutils.samples import f_2d, f_3d, generate_sample_data  # noqa: E402
from oracle import numpy_oracle as orc  # noqa: E402

import scipy  # noqa: E402

VERS = dict(numpy=np.__version__, scipy=scipy.__version__, python=sys.version.split()[0])


def versions():
    return np.array([f'{k}={v}' for k, v in VERS.items()])


def fb(x):
    """Second C2 response: f_3d + 0.5 sin(3 x0 + 1) sin(2.3 x1)."""
    return f_3d(x) + 0.5 * math.sin(3 * x[0] + 1) * math.sin(2.3 * x[1])


def ref_nll(krig, params):
    krig.records = None
    return krig._calc_likelihood_opti_exp(np.array(params, dtype=np.float64))


def ref_optimize(krig, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        krig.optimize(opti_algo='grid', **kw)


def conds(X, e, q):
    return np.array([orc.cond2(orc.corr_matrix(X, 10. ** e[c], q[c])) for c in range(e.shape[0])])


def save(name, **kw):
    path = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(path, versions=versions(), **kw)
    print('wrote', path, os.path.getsize(path), 'bytes', flush=True)


def gen_c1():
    """kriging_r2 setup (examples/kriging_r2.py:61-69,77-97,108-114,140)."""
    X = np.array(StructuredSample().generate_sample_plan(8, 1, [(0., 10.)]), dtype=np.float64)
    y = np.array([f_2d(x[0]) for x in X])
    k = Kriging(X, y)
    k.update_param([0.056389969498335794], [1.999966138140631])
    known_nll = k.calc_likelihood()
    known_pred = np.array([k.predict([2.0]), k.predict([6.0])])
    thetas = np.logspace(-2, 3, num=500)
    ps = np.linspace(1., 2., num=100)
    sweep = np.zeros((len(ps), len(thetas)))
    for it in range(len(thetas)):
        for ip in range(len(ps)):
            k.update_param([thetas[it]], [ps[ip]])
            sweep[ip][it] = k.calc_likelihood()
    grid_e, grid_p = orc.grid_candidates(1, 8, 5)
    grid_nll = np.array([ref_nll(k, np.concatenate([grid_e[c], grid_p[c]])) for c in range(40)])
    guess = LikeliOptimizer().generate_grid(lambda prm: ref_nll(k, prm), 1, 8, 5)
    ref_optimize(k, record_data=True)
    xq = np.linspace(0, 10, 1001)
    pred = np.array([k.predict([x]) for x in xq])
    save('c1_kriging_r2', X=X, y=y, known_theta=0.056389969498335794, known_p=1.999966138140631,
         known_nll=known_nll, known_pred=known_pred, sweep_theta=thetas, sweep_p=ps, sweep_nll=sweep,
         grid_nll=grid_nll, grid_guess=np.array(guess), opt_theta=k._theta, opt_p=k._p, opt_mu=k._mu,
         opt_nll=k.calc_likelihood(), n_records=len(k.records), xq=xq, pred=pred)
    # Halton-8 variant (BASELINE config text says "Halton samples")
    Xh = np.array(Halton().generate_sample_plan(8, 1, [(0., 10.)]), dtype=np.float64)
    yh = np.array([f_2d(x[0]) for x in Xh])
    kh = Kriging(Xh, yh)
    gh = np.array([ref_nll(kh, np.concatenate([grid_e[c], grid_p[c]])) for c in range(40)])
    save('c1_halton8', X=Xh, y=yh, grid_nll=gh)


def gen_c2_default():
    X = np.array(Halton().generate_sample_plan(64, 2, [(0., 10.), (0., 10.)]), dtype=np.float64)
    e, q = orc.grid_candidates(2, 8, 5)
    out = dict(X=X, grid_e=e, grid_p=q)
    for name, f in (('f3d', f_3d), ('fb', fb)):
        y = np.array([f(x) for x in X])
        k = Kriging(X, y)
        nll = np.array([ref_nll(k, np.concatenate([e[c], q[c]])) for c in range(e.shape[0])])
        out['y_' + name] = y
        out['nll_' + name] = nll
    out['cond'] = conds(X, e, q)
    # fitted model + predictions for the f3d response at the default-grid optimum (after SLSQP)
    y = out['y_f3d']
    k = Kriging(X, y)
    ref_optimize(k)
    rng = np.random.default_rng(7)
    xq = rng.uniform(0, 10, (256, 2))
    out.update(opt_theta=k._theta, opt_p=k._p, opt_mu=k._mu, opt_nll=k.calc_likelihood(), xq=xq,
               pred=np.array([k.predict(x) for x in xq]),
               opt_cond=orc.cond2(orc.corr_matrix(X, k._theta, k._p)))
    save('c2_halton64_default_grid', **out)


def _c2_chunk(args):
    name, lo, hi = args
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    X = np.array(Halton().generate_sample_plan(64, 2, [(0., 10.), (0., 10.)]), dtype=np.float64)
    f = f_3d if name == 'f3d' else fb
    y = np.array([f(x) for x in X])
    k = Kriging(X, y)
    e, q = orc.grid_candidates(2, 40, 25, lo, hi)
    return np.array([ref_nll(k, np.concatenate([e[c], q[c]])) for c in range(hi - lo)])


def gen_c2_full(procs):
    """Both full C2 grids (generate_grid(func, 2, 40, 25) = 1e6 candidates) on the reference.
    The 8 MB vectors are NOT committed; a stratified subset + the top candidates + every non-finite
    index are (tests/golden/c2_full_<resp>.npz)."""
    import multiprocessing as mp
    X = np.array(Halton().generate_sample_plan(64, 2, [(0., 10.), (0., 10.)]), dtype=np.float64)
    C, step = 1000000, 2500
    for name in ('f3d', 'fb'):
        t0 = time.time()
        with mp.Pool(procs) as pool:
            parts = pool.map(_c2_chunk, [(name, lo, lo + step) for lo in range(0, C, step)])
        nll = np.concatenate(parts)
        np.save(os.path.join('/tmp', f'c2_full_{name}_nll.npy'), nll)
        secs = time.time() - t0
        finite = np.isfinite(nll)
        order = np.argsort(np.where(finite, nll, np.inf), kind='stable')
        top = order[:256]
        strat = np.arange(0, C, 97)
        idx = np.unique(np.concatenate([top, strat]))
        e, q = orc.grid_candidates(2, 40, 25)
        save(f'c2_full_{name}', X=X, idx=idx, nll=nll[idx], cond=conds(X, e[idx], q[idx]),
             top_idx=top, nonfinite_idx=np.flatnonzero(~finite), argmin=int(orc.first_min_vec(nll)),
             ref_seconds=secs, ref_procs=procs)


def gen_c2_illcond():
    """The ill-conditioned corner of both full C2 grids (App. E rule iii), tests/golden/c2_illcond_<resp>.npz.
    S = candidates whose smallest fp64 Cholesky pivot is <= 1e-6 (3 971 of 1e6; R does not depend on y).  For every c in S:
      ref      the unmodified reference (read from the full run of gen_c2_full when /tmp still holds it, else recomputed)
      chol     oracle.nll_chol = np.linalg.cholesky (LAPACK potrf) formulation, +inf where potrf fails
      rounded  binary128 likelihood of the fp64 matrix the reference holds (oracle/quad_oracle.c), +inf = not PD
      exact    binary128 likelihood from the fp64 inputs (entries in 113 bits too)
      truth    100-digit mpmath Cholesky for the candidates whose fp64 matrix is not PD (what rule iii adjudicates with)
    """
    from oracle import quad, mp_oracle
    X = np.array(Halton().generate_sample_plan(64, 2, [(0., 10.), (0., 10.)]), dtype=np.float64)
    e, q = orc.grid_candidates(2, 40, 25)
    _, mp64 = quad.nll_f64(X, np.zeros(64) + np.arange(64), 10. ** e, q)
    S = np.flatnonzero(~(mp64 > 1e-6))
    print('ill-conditioned set:', len(S), flush=True)
    th = np.zeros_like(e)
    th[S] = [[10. ** v for v in row] for row in e[S]]  # scalar pow, as the reference forms theta (kriging.py:133-136)
    for name, f in (('f3d', f_3d), ('fb', fb)):
        y = np.array([f(x) for x in X])
        full = os.path.join('/tmp', f'c2_full_{name}_nll.npy')
        if os.path.exists(full):
            ref = np.load(full)[S]
        else:
            k = Kriging(X, y)
            ref = np.array([ref_nll(k, np.concatenate([e[c], q[c]])) for c in S])
        chol = np.array([orc.nll_chol(X, y, th[c], q[c]) for c in S])
        rounded, rmp = quad.nll_rounded(X, y, th[S], q[S])
        exact, emp = quad.nll_exact(X, y, th[S], q[S])
        truth = np.full(len(S), np.nan)
        for i in np.flatnonzero(~np.isfinite(rounded)):
            truth[i] = mp_oracle.nll_chol(X, y, th[S[i]], q[S[i]])
        save(f'c2_illcond_{name}', idx=S, ref=ref, chol=chol, rounded=rounded, rounded_min_pivot=rmp, exact=exact,
             exact_min_pivot=emp, truth=truth, f64_min_pivot=mp64[S])


def gen_c3():
    from scipy.stats import qmc
    X = qmc.LatinHypercube(d=8, seed=0).random(1024)
    y = np.sin(3 * X.sum(axis=1)) + X[:, 0] ** 2
    rng = np.random.default_rng(0)
    C = 100000
    e = rng.uniform(-3, 1, (C, 8))
    q = rng.uniform(1, 2, (C, 8))
    sel = np.array([0, 1, 2, 3, 4, 5, 6, 7, 50000, 99999])
    k = Kriging(X, y)
    t0 = time.time()
    nll = np.array([ref_nll(k, np.concatenate([e[c], q[c]])) for c in sel])
    secs = (time.time() - t0) / len(sel)
    # leave the model at candidate sel[0] and predict a few points (5 ms each)
    k.update_param(list(10. ** e[0]), q[0])
    xq = np.random.default_rng(3).uniform(0, 1, (32, 8))
    pred = np.array([k.predict(x) for x in xq])
    save('c3_lhs1024', sel=sel, e_sel=e[sel], p_sel=q[sel], nll=nll, cond=conds(X, e[sel], q[sel]),
         ref_seconds_per_eval=secs, xq=xq, pred=pred, X_checksum=float(X.sum()), y_checksum=float(y.sum()))
    # n = 256 mid-size case (cluster/blocked regime boundary), k = 8 and k = 2
    for kk in (2, 8):
        Xm = qmc.LatinHypercube(d=kk, seed=1).random(256)
        ym = np.sin(3 * Xm.sum(axis=1)) + Xm[:, 0] ** 2
        em = np.random.default_rng(5).uniform(-2, 1, (24, kk))
        qm = np.random.default_rng(6).uniform(1, 2, (24, kk))
        km = Kriging(Xm, ym)
        nl = np.array([ref_nll(km, np.concatenate([em[c], qm[c]])) for c in range(24)])
        save(f'mid_lhs256_k{kk}', X=Xm, y=ym, e=em, p=qm, nll=nl, cond=conds(Xm, em, qm))


def gen_c4():
    from scipy.stats import qmc
    out = {}
    # all five kernels on the kriging_r3 sample set (n = 36, k = 2)
    e6 = [0., 2., 4., 6., 8., 10.]
    px, py, pz = generate_sample_data(f_3d, e6, e6)
    X = np.array([px, py]).T
    y = np.array(pz)
    xq = np.random.default_rng(11).uniform(-2, 12, (64, 2))
    for name, a in (('lin', 0.17), ('cubic', 1.0), ('gaus', 1.0), ('mq', 0.5), ('imq', 1.3)):
        r = RBF(X, y)
        r.update_param(a, name)
        out[f'coeff_{name}'] = np.array(r.get_coeff())
        out[f'pred_{name}'] = np.array([r.predict(x) for x in xq])
        out[f'a_{name}'] = a
    save('rbf_r3_all_kernels', X=X, y=y, xq=xq, **out)
    # C4: Halton-4096 in 3-D, gaussian kernel (reference fit takes ~25 s)
    Xh = qmc.Halton(d=3, scramble=False).random(4097)[1:] * 10
    yh = np.sin(Xh[:, 0]) + 0.05 * Xh[:, 1] ** 2 - 0.01 * Xh[:, 2] ** 3
    t0 = time.time()
    r = RBF(Xh, yh)
    r.update_param(1.0, 'gaus')
    fit_s = time.time() - t0
    xq = np.random.default_rng(1).uniform(0, 10, (64, 3))
    t0 = time.time()
    pred = np.array([r.predict(x) for x in xq])
    pred_s = (time.time() - t0) / 64
    save('c4_rbf4096', coeff=np.array(r.get_coeff()), xq=xq, pred=pred, fit_seconds=fit_s,
         pred_seconds=pred_s, X_checksum=float(Xh.sum()), y_checksum=float(yh.sum()),
         cond=orc.cond2(orc.rbf_matrix(Xh, 2, 1.0)))


def gen_c5():
    e6 = [0., 2., 4., 6., 8., 10.]
    px, py, pz = generate_sample_data(f_3d, e6, e6)
    X = np.array([px, py]).T
    y = np.array(pz)
    k = Kriging(X, y)
    k.update_param([1e-3, 1e-3], [1.8, 1.8])
    known = k.calc_likelihood()
    e, q = orc.grid_candidates(2, 8, 5)
    nll = np.array([ref_nll(k, np.concatenate([e[c], q[c]])) for c in range(e.shape[0])])
    rng = np.random.default_rng(5)
    xq = np.concatenate([rng.uniform(-2, 12, (500, 2)), X[:8], [[1., 1.], [5.5, 3.3], [-2., 12.]]])
    k.update_param([0.05, 0.05], [1.9, 1.9])
    pred_fixed = np.array([k.predict(x) for x in xq])
    mu_fixed = k._mu
    kk = Kriging(X, y)
    ref_optimize(kk, record_data=True)
    pred_opt = np.array([kk.predict(x) for x in xq])
    save('c5_kriging_r3', X=X, y=y, known_nll=known, grid_nll=nll, grid_cond=conds(X, e, q), xq=xq,
         pred_fixed=pred_fixed, mu_fixed=mu_fixed, opt_theta=kk._theta, opt_p=kk._p, opt_mu=kk._mu,
         opt_nll=kk.calc_likelihood(), n_records=len(kk.records), pred_opt=pred_opt,
         opt_cond=orc.cond2(orc.corr_matrix(X, kk._theta, kk._p)))


def gen_validation():
    """mylibs/validation.py of the reference on the kriging_r3 data: fixed-parameter Kriging for the error
    measures, gaussian RBF for PRESS (its train() is a no-op, so PRESS is deterministic)."""
    from mylibs.validation import Validation
    e6 = [0., 2., 4., 6., 8., 10.]
    px, py, pz = generate_sample_data(f_3d, e6, e6)
    X = np.array([px, py]).T
    y = np.array(pz)
    rng = np.random.default_rng(21)
    vx = rng.uniform(0, 10, (40, 2))
    vy = np.array([f_3d(v) for v in vx])
    k = Kriging(X, y)
    k.update_param([0.05, 0.05], [1.9, 1.9])
    v = Validation()
    r = RBF(X, y)
    r.update_param(1.0, 'gaus')
    save('validation_r3', X=X, y=y, vx=vx, vy=vy,
         krig_deviation=v.calc_deviation(vx, vy, k.predict), krig_rmse=v.calc_rmse(vx, vy, k.predict),
         krig_mae=v.calc_mae(vx, vy, k.predict), krig_rae=v.calc_rae(vx, vy, k.predict),
         rbf_rmse=v.calc_rmse(vx, vy, r.predict), rbf_press=v.calc_press(X, y, r.predict, RBF, update_params=[1.0, 'gaus']),
         halton_7x3=np.array(Halton().generate_sample_plan(7, 3, [(0., 1.), (-1., 1.), (2., 5.)])),
         structured_9x2=np.array(StructuredSample().generate_sample_plan(9, 2, [(0., 10.), (-5., 5.)])))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--full-c2', action='store_true')
    ap.add_argument('--only', default='')
    ap.add_argument('--procs', type=int, default=os.cpu_count())
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    todo = dict(c1=gen_c1, c2=gen_c2_default, c3=gen_c3, c4=gen_c4, c5=gen_c5, val=gen_validation)
    if a.only == 'illcond':
        gen_c2_illcond()
    elif a.full_c2:
        gen_c2_full(a.procs)
    else:
        for name, fn in todo.items():
            if not a.only or name in a.only.split(','):
                t0 = time.time()
                fn()
                print(name, 'done in', round(time.time() - t0, 1), 's', flush=True)
