"""Pins the CPU oracle (oracle/numpy_oracle.py) against fixtures produced by executing the UNMODIFIED
reference (oracle/gen_golden.py).  The reference ships no tests of its own; its single
in-repo known answer (examples/kriging_r2.py:77-79) is part of the C1 fixture."""
import numpy as np
import pytest

from oracle import numpy_oracle as orc


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.abs(a - b) / np.maximum(1, np.abs(b))


def tol(cond):
    """Conditioning-aware parity tolerance (DESIGN.md 'Numerical contract'): 1e-9 relative for
    cond2(R) <= 5e6, growing linearly with cond beyond.  Measured: at cond 5.6e7 the reference is
    +3.1e-9 and the Cholesky forms -1.8e-9 .. -2.9e-9 away from the 60-digit mpmath value (candidate 1481
    of the default grid, response fb), i.e. they differ from EACH OTHER by up to 6e-9 there."""
    return 1e-9 * np.maximum(1.0, np.asarray(cond, dtype=float) / 5e6)


def test_known_answer_kriging_r2(gold):
    g = gold('c1_kriging_r2')
    X, y = g['X'], g['y']
    th, p = [float(g['known_theta'])], [float(g['known_p'])]
    # -0.5010025589591347
    assert abs(float(g['known_nll']) - (-0.5010025589591347)) < 1e-12
    assert orc.nll_lu_loops(X, y, th, p) == float(g['known_nll'])       # bit-faithful restatement
    assert rel(orc.nll_lu(X, y, th, p), g['known_nll']) < 1e-12
    assert rel(orc.nll_chol(X, y, th, p), g['known_nll']) < 1e-10
    pred = orc.predict_lu(X, y, th, p, np.array([[2.0], [6.0]]))
    assert np.all(rel(pred, g['known_pred']) < 1e-11)


def test_c1_sweep_lu_and_chol(gold):
    g = gold('c1_kriging_r2')
    X, y = g['X'], g['y']
    ths, ps, ref = g['sweep_theta'], g['sweep_p'], g['sweep_nll']
    ip = np.arange(0, 100, 9)
    it = np.arange(0, 500, 23)
    for a in ip:
        for b in it:
            lu = orc.nll_lu(X, y, [ths[b]], [ps[a]])
            assert rel(lu, ref[a, b]) < 1e-9 or orc.cond2(orc.corr_matrix(X, [ths[b]], [ps[a]])) > 1e8
    # reference argmin of the sweep: (ip, it) = (99, 75)
    assert np.unravel_index(np.argmin(ref), ref.shape) == (99, 75)
    ch = orc.nll_chol(X, y, [ths[75]], [ps[99]])
    assert rel(ch, ref[99, 75]) < 1e-9


def test_grid_enumeration_and_first_min(gold):
    g = gold('c5_kriging_r3')
    e, q = orc.grid_candidates(2, 8, 5)
    # App. C check: c = 960 -> [-5, -5, 1.0, 1.75]
    assert list(e[960]) == [-5.0, -5.0] and list(q[960]) == [1.0, 1.75]
    assert list(e[1]) == [np.linspace(-5, 5, 8)[1], -5.0]            # dimension 0 is the fastest digit
    nll = g['grid_nll']
    assert orc.first_min(nll) == orc.first_min_vec(nll) == 960
    assert orc.first_min([np.nan, 3.0, 1.0, 1.0, np.inf]) == 2        # first strict minimum, NaN never wins
    assert orc.first_min([np.inf, np.nan]) == -1
    lo, hi = 700, 731
    e2, q2 = orc.grid_candidates(2, 8, 5, lo, hi)
    assert np.array_equal(e2, e[lo:hi]) and np.array_equal(q2, q[lo:hi])


def test_default_grid_values_n36(gold):
    g = gold('c5_kriging_r3')
    X, y = g['X'], g['y']
    e, q = orc.grid_candidates(2, 8, 5)
    ref, cond = g['grid_nll'], g['grid_cond']
    assert rel(orc.nll_lu(X, y, [1e-3, 1e-3], [1.8, 1.8]), g['known_nll']) < 1e-11
    assert abs(float(g['known_nll']) - (-53.655056991209705)) < 1e-9
    sel = np.arange(0, 1600, 13)
    for c in sel:
        if cond[c] <= 1e8 and np.isfinite(ref[c]):
            assert rel(orc.nll_lu(X, y, 10. ** e[c], q[c]), ref[c]) < tol(cond[c])
            assert rel(orc.nll_chol(X, y, 10. ** e[c], q[c]), ref[c]) < tol(cond[c])
    assert np.sum(~np.isfinite(ref)) == 14


def test_default_grid_n64_argmin_and_values(gold):
    g = gold('c2_halton64_default_grid')
    X = g['X']
    e, q = g['grid_e'], g['grid_p']
    assert orc.first_min_vec(g['nll_f3d']) == 1489
    assert orc.first_min_vec(g['nll_fb']) == 1297                     # SURVEY.md App. E counter-check
    cond = g['cond']
    for name in ('f3d', 'fb'):
        y, ref = g['y_' + name], g['nll_' + name]
        for c in range(5, 1600, 41):
            if cond[c] <= 1e8 and np.isfinite(ref[c]):
                assert rel(orc.nll_chol(X, y, 10. ** e[c], q[c]), ref[c]) < tol(cond[c])


def test_predictions(gold):
    g = gold('c5_kriging_r3')
    X, y, xq = g['X'], g['y'], g['xq']
    th, p = [0.05, 0.05], [1.9, 1.9]
    assert np.all(rel(orc.predict_lu(X, y, th, p, xq), g['pred_fixed']) < 1e-10)
    assert np.all(rel(orc.predict_chol(X, y, th, p, xq), g['pred_fixed']) < 1e-9)
    assert rel(orc.fit_chol(X, y, th, p)['mu'], g['mu_fixed']) < 1e-10


def test_mse_against_mpmath(gold):
    """MSE is not in the reference ('parity unpinned'): pinned against 60-digit mpmath instead."""
    g = gold('mp_kriging_r3')
    X, y, xq = g['X'], g['y'], g['xq']
    for tag in ('fixed', 'rough'):
        th, p = g['theta_' + tag], g['p_' + tag]
        yy, mse = orc.predict_chol(X, y, th, p, xq, return_mse=True)
        s2 = float(g['s2_' + tag])
        assert np.all(rel(yy, g['y_' + tag]) < 1e-9)
        assert np.max(np.abs(mse - np.maximum(g['mse_' + tag], 0))) < 1e-9 * s2
        assert rel(orc.nll_chol(X, y, th, p), g['nll_' + tag]) < 1e-10
    # SURVEY.md §8c values for the 'fixed' fit
    assert abs(g['mse_fixed'][-3] - 0.29801489720680) < 1e-9
    assert abs(g['mse_fixed'][-2] - 0.180151619545) < 1e-9
    assert abs(g['mse_fixed'][-1] - 6.1652000483847) < 1e-9


def test_rbf_all_kernels(gold):
    g = gold('rbf_r3_all_kernels')
    X, y, xq = g['X'], g['y'], g['xq']
    for name in ('lin', 'cubic', 'gaus', 'mq', 'imq'):
        kind, a = orc.RBF_KINDS[name], float(g['a_' + name])
        c = orc.rbf_fit(X, y, kind, a)
        assert np.max(np.abs(c - g['coeff_' + name])) <= 1e-9 * np.max(np.abs(g['coeff_' + name]))
        assert np.all(rel(orc.rbf_predict(X, c, kind, a, xq), g['pred_' + name]) < 1e-9)


@pytest.mark.parametrize('kk', [2, 8])
def test_mid_n256(gold, kk):
    g = gold(f'mid_lhs256_k{kk}')
    X, y = g['X'], g['y']
    for c in range(0, 24, 5):
        if g['cond'][c] <= 1e8:
            assert rel(orc.nll_chol(X, y, 10. ** g['e'][c], g['p'][c]), g['nll'][c]) < tol(g['cond'][c])


# ---------------------------------------------------------------------------------------------- extended precision
def test_quad_oracle_against_mpmath(gold):
    """oracle/quad_oracle.c (binary128) == 100-digit mpmath Cholesky, well- and ill-conditioned (kriging_r3 data)."""
    from oracle import quad, mp_oracle
    g = gold('c5_kriging_r3')
    X, y = g['X'], g['y']
    for th, p in (([0.05, 0.05], [1.9, 1.9]), ([1e-3, 1e-3], [1.8, 1.8]), ([1e-4, 3e-4], [2.0, 1.95])):
        q, _ = quad.nll_exact(X, y, [th], [p])
        t = mp_oracle.nll_chol(X, y, th, p)
        assert abs(q[0] - t) <= 1e-12 * max(1, abs(t)), (th, q[0], t)
    # candidate 998813 of the full C2 grid, exact value -264.6644
    ic = gold('c2_illcond_f3d')
    i = int(np.flatnonzero(ic['idx'] == 998813)[0])
    assert abs(ic['exact'][i] - (-264.6644)) < 1e-3 and abs(ic['ref'][i] - (-268.1161655258276)) < 1e-9


def test_c2_illcond_fixture(gold):
    """The ill-conditioned corner of the C2 grids (oracle/gen_golden.py --only illcond): LAPACK's Cholesky accepts
    only matrices that ARE positive definite (binary128), the argmin over 'fp64 matrix is PD' is the reference's own
    998813 on f_3d and 960447 on fb, and every reference value that undercuts it on a non-PD matrix is an artefact."""
    from oracle import quad
    from sm_b200 import workloads
    for resp, want in (('f3d', 998813), ('fb', 960447)):
        ic, g = gold(f'c2_illcond_{resp}'), gold(f'c2_full_{resp}')
        assert not (np.isfinite(ic['chol']) & ~np.isfinite(ic['rounded'])).any()
        pd = np.isfinite(ic['rounded'])
        best = ic['rounded'][pd].min()
        if resp == 'f3d':
            assert int(ic['idx'][pd][np.argmin(ic['rounded'][pd])]) == want == int(g['argmin'])
        else:
            assert best > -41.300450780839526                        # nothing in the corner beats the genuine optimum
            best = -41.300450780839526
        under = np.isfinite(ic['ref']) & (ic['ref'] < best) & ~pd
        assert under.sum() >= 1 and np.all(ic['truth'][under] > best)
        # the fixture is reproducible from the committed oracle
        X, y = workloads.c2_halton64(resp)
        e, q = orc.grid_candidates(2, 40, 25)
        sel = np.arange(0, len(ic['idx']), 97)
        th = np.array([[10. ** v for v in row] for row in e[ic['idx'][sel]]])
        r, _ = quad.nll_rounded(X, y, th, q[ic['idx'][sel]])
        assert np.array_equal(r, ic['rounded'][sel])


def test_product_double_double_header_on_host(gold, tmp_path):
    """csrc/dd.cuh (the arithmetic of the GPU's extended-precision pass) compiled for the host: exp / pow correctly
    rounded (checked against mpmath), and the double-double likelihood accepts exactly the candidates whose fp64 matrix
    is positive definite in binary128, with the same value."""
    import ctypes
    import os
    import subprocess
    import mpmath as mp
    from sm_b200 import workloads
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / 'libdd_check.so')
    subprocess.run(['g++', '-O2', '-ffp-contract=off', '-fopenmp', '-shared', '-fPIC',
                    os.path.join(root, 'tools', 'dd_check.cpp'), '-o', so], check=True)
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    P = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    rng = np.random.default_rng(0)
    x = -rng.uniform(0, 40, 3000) * rng.uniform(0, 1, 3000) ** 2
    out = np.empty_like(x)
    lib.exp_cr_vec(P(x), len(x), P(out))
    mp.mp.prec = 200
    assert all(o == float(mp.exp(mp.mpf(float(v)))) for o, v in zip(out, x))
    xs, ps = rng.uniform(0, 10, 3000), rng.uniform(1, 2, 3000)
    lib.pow_cr_vec(P(xs), P(ps), len(xs), P(out))
    assert all(o == float(mp.power(mp.mpf(float(a)), mp.mpf(float(b)))) for o, a, b in zip(out, xs, ps))
    ic = gold('c2_illcond_f3d')
    X, y = workloads.c2_halton64('f3d')
    e, q = orc.grid_candidates(2, 40, 25)
    sel = np.flatnonzero(ic['f64_min_pivot'] <= 1e-9)
    th = np.ascontiguousarray([[10. ** v for v in row] for row in e[ic['idx'][sel]]])
    qq = np.ascontiguousarray(q[ic['idx'][sel]])
    nll, mpv = np.empty(len(sel)), np.empty(len(sel))
    lib.dd_nll_batch(P(np.ascontiguousarray(X)), P(np.ascontiguousarray(y)), 64, 2, P(th), P(qq), ctypes.c_long(len(sel)),
                     P(nll), P(mpv))
    r, rmp = ic['rounded'][sel], ic['rounded_min_pivot'][sel]
    assert np.array_equal(np.isfinite(nll), np.isfinite(r))
    m = np.isfinite(r)
    assert np.all(np.abs(nll[m] - r[m]) <= 1e-9 * np.maximum(1, np.abs(r[m])) + 1e-13 / rmp[m])
    assert int(ic['idx'][sel][np.argmin(nll)]) == 998813
