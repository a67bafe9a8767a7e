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
