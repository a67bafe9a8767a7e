"""n-dimensional ordinary Kriging on one B200 -- drop-in for the reference's ``mylibs/kriging.py``.

Same class, methods, attributes and error behaviour as the reference (file:line cited per method); the
arithmetic runs in hand-written sm_100a kernels behind the C-ABI of ``include/surrogate_b200.h``.  There
is no CPU path: without a CUDA device every compute call raises ``SurrogateB200Error``.

Differences that are deliberate (DESIGN.md "Numerical contract"):
  * the likelihood is evaluated in Cholesky form; a correlation matrix that is not numerically SPD gives
    ``+inf`` where the reference (LU inverse + log|det|, kriging.py:79,100) returns rounding noise;
  * ``predict`` uses weights cached at ``update_param`` time instead of recomputing ``psi^T R^-1`` per call;
  * additive batched entry points: ``likelihood_batch``, ``likelihood_grid``, ``grid_search``,
    ``predict_batch`` (value and MSE).
"""
import numpy as np
from scipy.optimize import basinhopping, minimize

from mylibs.likeli_optimizer import LikeliOptimizer
from myutils.time_track import TimeTrack
from sm_b200 import _cabi, device, dist

VERBOSE = False
import os
LAZY_FIT_MIN_N = int(os.environ.get('SMB_LAZY_FIT_MIN_N', '145'))  # from here on the optimiser objective is one single-candidate launch of the batched kernel


class Kriging:

    def __init__(self, known_in, known_val):
        """
        :param known_in: list of lists with input sample points
        :param known_val: list of results for the known_in
        (reference: mylibs/kriging.py:28-44)
        """
        self._known_in = np.array(known_in, dtype=np.float64)
        self._known_val = np.array(known_val, dtype=np.float64)
        if len(self._known_in.shape) == 1:
            self._known_in = self._known_in.reshape((self._known_in.shape[0], 1))
        self._k = self._known_in.shape[1]
        self._n = self._known_in.shape[0]
        self._theta = 1. * np.ones((self._k, 1)).flatten()
        self._p = 2. * np.ones((self._k, 1)).flatten()
        self._mu = None
        self._mat_u = None
        # device state (created on first use)
        self._d = None
        self._fit = None
        self._fit_stale = False  # parameters changed through the optimiser objective: fit deferred (n >= LAZY_FIT_MIN_N)
        self._cor_mat_cache = None
        self._cor_mat_inv_cache = None

    # ------------------------------------------------------------------ device plumbing
    def _dev(self):
        if self._d is None:
            device.require_cuda()
            self._d = dict(X=device.to_dev(self._known_in), y=device.to_dev(self._known_val))
        return self._d

    def _run_fit(self, want_R=False, want_mse=True):
        """sm_krig_fit at the current (theta, p): the state update_param leaves behind.  want_mse=False at
        n >= LAZY_FIT_MIN_N skips L^-1 (only the MSE needs it): the fit is then the dataflow likelihood kernel on one
        candidate plus a back-substitution (5 ms instead of 19 ms at n = 1024)."""
        L = _cabi.lib()
        d = self._dev()
        n, k = self._n, self._k
        npad = L.sm_krig_npad(n)
        th, pw = device.to_dev(self._theta), device.to_dev(self._p)
        if th.numel() != k or pw.numel() != k:
            raise ValueError('theta and p need one entry per input dimension')
        with_mop = want_mse or n < LAZY_FIT_MIN_N
        f = dict(theta=th, p=pw, scalars=device.empty(4), w=device.empty(n), b=device.empty(npad),
                 rres=device.empty(npad), Mop=device.empty(L.sm_krig_mop_doubles(n)) if with_mop else None,
                 R=device.empty((n, n)) if want_R else None)
        _cabi.check(L.sm_krig_fit(device.ptr(d['X']), device.ptr(d['y']), n, k, device.ptr(th), device.ptr(pw),
                                  device.ptr(f['scalars']), device.ptr(f['w']), device.ptr(f['b']),
                                  device.ptr(f['rres']), device.ptr(f['Mop']), device.ptr(f['R']),
                                  device.stream()))
        sc = f['scalars'].cpu().numpy()  # synchronises
        # info (sc[3]): 0 fp64 fit, 2 / 3 ill-conditioned -> values from the double-double factorisation (3: no MSE operand),
        # 1 not positive definite
        f['nll'], f['mu'], f['sigma2'], f['ok'], f['info'] = float(sc[0]), float(sc[1]), float(sc[2]), sc[3] != 1.0, int(sc[3])
        return f

    def _ensure_mse_state(self):
        """L^-1 (the MSE operand) is computed on demand when the last fit was weights-only."""
        self._ensure_fit()
        if self._fit['Mop'] is None:
            self._fit = self._run_fit(want_mse=True)

    # ------------------------------------------------------------------ reference API
    def train(self):
        """trains the surrogate (reference: kriging.py:46-51)"""
        self.optimize()

    def update_param(self, theta, p):
        """
        updates the parameters of the surrogate model (reference: kriging.py:53-63)
        :param theta: vector of theta parameters for each entry one (from range [1e-5 .. 1e+5])
        :param p: vector of p parameters for each entry one (from range[1 ..2])
        """
        self._theta = np.array(theta, dtype=np.float64)
        self._p = np.array(p, dtype=np.float64)
        self._calc_cormat()
        self._calc_mu()

    def _calc_cormat(self):
        """Builds R on the device, factorises it and caches the fitted state (reference: kriging.py:66-84).
        Returns nothing heavy: the dense matrix is materialised lazily through ``_cor_mat``."""
        self._cor_mat_cache = None
        self._cor_mat_inv_cache = None
        self._fit = self._run_fit(want_mse=False)
        self._fit_stale = False
        if not self._fit['ok'] and VERBOSE:
            print('ERROR: correlation matrix is not numerically positive definite')

    def _ensure_fit(self):
        """(Re)runs the fit when there is none yet or the optimiser objective left it behind the parameters."""
        if self._fit is None or self._fit_stale:
            self._calc_cormat()
            self._mu = np.float64(self._fit['mu'])

    def _objective_nll(self, theta, p):
        """update_param + calc_likelihood for the optimiser objectives (kriging.py:118-141).  For n >= LAZY_FIT_MIN_N
        the likelihood comes from ONE single-candidate launch of the batched kernel and the full fit (weights,
        L^-1 for the MSE: 18 ms at n = 1024) is deferred until something reads the fitted state -- SLSQP makes ~100
        sequential objective calls.  Smaller n keeps the fit-per-call path (tens of microseconds)."""
        if self._n < LAZY_FIT_MIN_N:
            self.update_param(theta, p)
            return self.calc_likelihood()
        self._theta = np.array(theta, dtype=np.float64)
        self._p = np.array(p, dtype=np.float64)
        self._cor_mat_cache = None
        self._cor_mat_inv_cache = None
        self._fit_stale = True
        nll = self._likelihood_list(self._theta.reshape(1, -1), self._p.reshape(1, -1))[0][0]
        return float('inf') if np.isnan(nll) else np.float64(nll)

    @property
    def batched_stencil(self):
        """LikeliOptimizer.find hands SciPy's finite-difference points of a gradient to _objective_map when the
        objective is a single-candidate launch anyway (n >= LAZY_FIT_MIN_N): 2k sequential launches become one.
        (At small n the per-call table set-up makes the batch slower than 2k fits -- measured)"""
        return self._n >= LAZY_FIT_MIN_N

    def _objective_map(self, fun, iterable):
        """map-like `workers` callable for scipy.optimize.minimize (SciPy >= 1.16): the values of
        _calc_likelihood_opti_exp at every point of `iterable`, from one batched launch.  `fun` (SciPy's wrapper of
        the objective) is only used when a point is not a plain [log10 theta, p] vector."""
        xs = [np.asarray(x, dtype=np.float64).ravel() for x in iterable]
        if not xs:
            return []
        if any(x.shape[0] != 2 * self._k or np.isnan(x).any() for x in xs):
            return [fun(x) for x in xs]
        theta = np.array([[10. ** e for e in x[:self._k]] for x in xs])  # scalar pow, like the objective
        p = np.array([x[self._k:] for x in xs])
        nll = self._likelihood_list(theta, p)[0]
        if getattr(self, 'records', None) is not None:
            self.records.extend(xs)
        # the object's parameters follow the last evaluated point, as after a sequential loop
        self._theta, self._p = theta[-1].copy(), p[-1].copy()
        self._cor_mat_cache = None
        self._cor_mat_inv_cache = None
        self._fit_stale = True
        return [float('inf') if np.isnan(v) else np.float64(v) for v in nll]

    def _calc_mu(self):
        """mu = 1^T R^-1 y / 1^T R^-1 1 (reference: kriging.py:86-93); computed by the fit kernel."""
        self._ensure_fit()
        self._mu = np.float64(self._fit['mu'])
        return self._mu

    @property
    def _cor_mat(self):
        """Dense correlation matrix (reference attribute, kriging.py:80); None before update_param."""
        if self._fit is None and not self._fit_stale:
            return None
        if self._cor_mat_cache is None:
            self._cor_mat_cache = self._run_fit(want_R=True)['R'].cpu().numpy()
        return self._cor_mat_cache

    @property
    def _core_mat_inv(self):
        """R^-1 = L^-T L^-1 assembled on the device from the cached factor (reference attribute,
        kriging.py:79).  Compatibility accessor, not on the hot path."""
        if self._fit is None and not self._fit_stale:
            return None
        self._ensure_mse_state()
        if self._cor_mat_inv_cache is None:
            import torch
            n = self._n
            nt = (n + 7) // 8
            mop = self._fit['Mop'].view(-1, 8, 8)  # operand order: [r][2 (c & 3) + (c >> 2)]
            perm = torch.tensor([0, 4, 1, 5, 2, 6, 3, 7], device=mop.device)  # position -> column
            tiles = torch.empty_like(mop)
            tiles[:, :, perm] = mop
            M = torch.zeros((nt * 8, nt * 8), dtype=torch.float64, device=mop.device)
            t = 0
            for i in range(nt):
                for j in range(i + 1):
                    M[8 * i:8 * i + 8, 8 * j:8 * j + 8] = tiles[t]
                    t += 1
            M = M[:n, :n]
            self._cor_mat_inv_cache = (M.T @ M).cpu().numpy()
        return self._cor_mat_inv_cache

    def calc_likelihood(self):
        """
        negative concentrated log-likelihood at the current parameters (reference: kriging.py:95-116)
        :return: neg. log. likelihood, or inf when R is not numerically SPD / sigma^2 is invalid
        """
        self._ensure_fit()
        nll = self._fit['nll']
        if np.isnan(nll):
            return float('inf')
        return np.float64(nll)

    def _calc_likelihood_opti_theta_only(self, params, *args):
        """(reference: kriging.py:118-121)"""
        return self._objective_nll(params, args[0])

    def _calc_likelihood_opti(self, params, *args):
        """(reference: kriging.py:123-128)"""
        neg_ln_like = self._objective_nll(params[0:self._k], params[self._k:])
        if getattr(self, 'records', None) is not None:
            self.records.append(params)
        return neg_ln_like

    def _calc_likelihood_opti_exp(self, params, *args):
        """objective of the optimisers: params = [log10 theta_1..k, p_1..k] (reference: kriging.py:130-141)"""
        if np.isnan(params).any():
            return float('nan')
        exps = params[0:self._k]
        thetas = []
        for e in exps:
            thetas.append(10.**e)
        neg_ln_like = self._objective_nll(thetas, params[self._k:])
        if getattr(self, 'records', None) is not None:
            self.records.append(params)
        return neg_ln_like

    def optimize_theta_only(self):
        """(reference: kriging.py:143-152)"""
        x0 = np.ones((self._k, 1)).flatten()
        bnds = [(0.0001, 1000.) for _ in range(self._k)]
        opt = {'disp': True, 'maxiter': 99999}
        res = minimize(self._calc_likelihood_opti_theta_only, x0, args=self._p, method='SLSQP', tol=1e-6,
                       options=opt, bounds=bnds)
        self._theta = res.x

    def optimize(self, init_guess=None, opti_algo='grid', record_data=False):
        """
        runs automatic optimization of thetas and ps (reference: kriging.py:154-204)
        :param init_guess: list of input values for an initial guess ('basin' only)
        :param opti_algo: 'grid' (LikeliOptimizer: batched GPU grid + SLSQP), 'basin' (scipy basin-hopping with the
                          reference's step / accept rules) or -- additive -- 'restarts' (seeded random starts, evaluated
                          as one batch, SLSQP from the best few; the deterministic replacement of 'basin')
        :param record_data: record every evaluated parameter vector in self.records
        """
        timer = TimeTrack('optiTimer')
        self.records = [] if record_data else None
        searches = (('basin', lambda: self._search_basin(init_guess)),
                    ('grid', lambda: LikeliOptimizer(debug=True).find(self._calc_likelihood_opti_exp, self._k)),
                    ('restarts', self._search_restarts))
        for key, search in searches:      # substring match in the reference's order (kriging.py:166,191)
            if key in opti_algo:
                timer.tic()
                res = search()
                timer.toc(print_it=True)
                break
        else:
            raise Exception('ERROR: unknown optimizer selected')
        if record_data:
            print('Kriging Likelihood optimization evaluations: {:d}'.format(len(self.records)))
        x = res.x
        self.update_param([10. ** e for e in x[0:self._k]], x[self._k:])

    def _search_box(self):
        """[-5, 5]^k for log10(theta), [1, 2]^k for p (kriging.py:173-177)"""
        return [(-5., +5.)] * self._k + [(1., 2.)] * self._k

    def _search_basin(self, init_guess):
        """scipy basin-hopping exactly as the reference configures it (kriging.py:166-190): 1000 hops, stop after 100
        without improvement, SLSQP inner minimiser, BasinHoppingStep proposals filtered by BasinHoppingBounds.  Draws
        from numpy's global generator like the reference (seed it with np.random.seed for a reproducible run)."""
        if init_guess is None:
            init_guess = list(self._theta) + list(self._p)
        box = self._search_box()
        lo, hi = zip(*box)
        inner = dict(method='SLSQP', bounds=box, options={'disp': False, 'maxiter': 5e3}, tol=1e-4)
        return basinhopping(self._calc_likelihood_opti_exp, init_guess, minimizer_kwargs=inner,
                            accept_test=BasinHoppingBounds(xmax=hi, xmin=lo), take_step=BasinHoppingStep(),
                            niter=1000, niter_success=100)

    restart_seed = 0       # 'restarts': generator seed, number of random starts per batch and of SLSQP refinements
    restart_points = 4096
    restart_refine = 4

    def _search_restarts(self):
        """Deterministic global search: `restart_points` uniformly random parameter vectors in
        the search box are evaluated in ONE likelihood_batch launch, SLSQP refines the best `restart_refine` of them,
        the best end point wins (ties: the earlier start)."""
        rng = np.random.default_rng(self.restart_seed)
        lo, hi = (np.array(v) for v in zip(*self._search_box()))
        starts = rng.uniform(lo, hi, (self.restart_points, 2 * self._k))
        nll = self.likelihood_batch(starts[:, :self._k], starts[:, self._k:])
        if self.records is not None:
            self.records.extend(starts)
        order = np.argsort(np.where(np.isfinite(nll), nll, np.inf), kind='stable')[:self.restart_refine]
        best = None
        for i in order:
            res = minimize(self._calc_likelihood_opti_exp, starts[i], method='SLSQP', tol=1e-8,
                           options={'disp': False, 'maxiter': 5e3}, bounds=self._search_box())
            if best is None or res.fun < best.fun:
                best = res
        return best

    def grid_top(self, theta_sections=8, p_sections=5, m=4):
        """parameter vectors of the m best (finite) grid candidates, best first, ties by index (for multi-start SLSQP)"""
        _, _, nll = self.grid_search(theta_sections, p_sections, return_nll=True)
        order = np.argsort(np.where(np.isfinite(nll), nll, np.inf), kind='stable')[:m]
        return [grid_param(self._k, theta_sections, p_sections, int(c)) for c in order if np.isfinite(nll[c])]

    def predict(self, x_pred):
        """
        predicts a value from the surrogate model (reference: kriging.py:206-220)
        :param x_pred: vector of input values (anything indexable with k entries)
        :return: result value (numpy float64 scalar, usable as a scipy objective)
        """
        xq = np.array([[x_pred[i] for i in range(self._k)]], dtype=np.float64)
        return self.predict_batch(xq)[0]

    # ------------------------------------------------------------------ batched entry points (additive)
    def likelihood_batch(self, log10_theta, p, return_best=False, distributed=None, gather=True):
        """nll for every row of an explicit candidate list: log10_theta[C,k], p[C,k] -> nll[C].
        Replaces a loop of _calc_likelihood_opti_exp calls (kriging.py:130-141).
        With torch.distributed initialised (and distributed is not False) the rows are sharded contiguously over the
        ranks (every rank holds the replicated samples); the per-rank (nll_min, row) pairs cross the
        interconnect as one 16-byte all_gather and, with gather=True, the nll slices are all-gathered so every rank
        returns the full vector (gather=False: this rank's rows [lo, hi) only).  Row indices are global."""
        e = np.atleast_2d(np.asarray(log10_theta, dtype=np.float64))
        q = np.atleast_2d(np.asarray(p, dtype=np.float64))
        if e.shape != q.shape or e.shape[1] != self._k:
            raise ValueError('candidate tables must both be (C, k)')
        C = e.shape[0]
        if C == 0:
            return (np.empty(0), float('inf'), -1) if return_best else np.empty(0)
        rank, world = dist.rank_world(distributed)
        lo, hi = dist.shard_range(C, rank, world)
        e_d = device.to_dev(e[lo:hi])
        th_d = device.empty(e_d.shape)  # theta = 10.**e on the device, correctly rounded like the reference's scalar pow
        _cabi.check(_cabi.lib().sm_exp10(device.ptr(e_d), e_d.numel(), device.ptr(th_d), device.stream()))
        nll_d, bv_d, bi_d = self._likelihood_list_dev(th_d, q[lo:hi], lo)
        if world > 1:
            bv, bi = dist.global_first_min_dev(bv_d, bi_d) if return_best else (None, None)
            if gather:
                nll_d, lo, hi = dist.gather_slices(nll_d, C), 0, C
        elif return_best:
            bv, bi = float(bv_d.item()), int(bi_d.item())
        out = nll_d.cpu().numpy()
        nanrow = np.isnan(e[lo:hi]).any(axis=1) | np.isnan(q[lo:hi]).any(axis=1)
        out[nanrow] = np.nan  # NaN parameters -> NaN (kriging.py:131-132)
        if return_best:
            return out, bv, bi
        return out

    def _likelihood_list_dev(self, theta_lin, p, idx_base=0):
        """sm_krig_nll_batch on linear theta[C,k], p[C,k]: device tensors (nll[C], best value, best index + idx_base)"""
        L = _cabi.lib()
        d = self._dev()
        C = theta_lin.shape[0]
        theta = device.to_dev(theta_lin)
        pd = device.to_dev(p)
        nll = device.empty(C)
        bv, bi = device.empty(1), device.empty(1, dtype=_int64())
        _cabi.check(L.sm_krig_nll_batch(device.ptr(d['X']), device.ptr(d['y']), self._n, self._k, device.ptr(theta),
                                        device.ptr(pd), C, int(idx_base), device.ptr(nll), device.ptr(bv), device.ptr(bi),
                                        device.stream()))
        return nll, bv, bi

    def _likelihood_list(self, theta_lin, p):
        """sm_krig_nll_batch on linear theta[C,k], p[C,k] -> (nll[C], best value, best index), on the host"""
        nll, bv, bi = self._likelihood_list_dev(theta_lin, p)
        return nll.cpu().numpy(), float(bv.item()), int(bi.item())

    def likelihood_grid(self, theta_sections=8, p_sections=5, c_begin=0, c_count=None, return_nll=True, as_device=False):
        """nll over candidates [c_begin, c_begin + c_count) of LikeliOptimizer.generate_grid's enumeration
        (likeli_optimizer.py:50-97) on THIS device.  Returns (best_val, best_idx, nll or None);
        best_idx is the global flat index of the first strict minimum, -1 if nothing is finite."""
        L = _cabi.lib()
        d = self._dev()
        T, P = int(theta_sections), int(p_sections)
        total = (T ** self._k) * (P ** self._k)
        c_count = total - c_begin if c_count is None else int(c_count)
        th_axis, p_axis = grid_axes(T, P)
        th_lin = device.to_dev(np.array([10.**e for e in th_axis]))  # scalar pow, like kriging.py:135-136
        pa = device.to_dev(p_axis)
        nll = device.empty(c_count) if return_nll else None
        bv, bi = device.empty(1), device.empty(1, dtype=_int64())
        _cabi.check(L.sm_krig_nll_grid(device.ptr(d['X']), device.ptr(d['y']), self._n, self._k, device.ptr(th_lin), T,
                                       device.ptr(pa), P, int(c_begin), c_count, device.ptr(nll), device.ptr(bv),
                                       device.ptr(bi), device.stream()))
        if as_device:  # no synchronisation: (1 x float64, 1 x int64, nll) stay on the device
            return bv, bi, nll
        return float(bv.item()), int(bi.item()), nll

    def grid_search(self, theta_sections=8, p_sections=5, distributed=None, return_nll=False):
        """The grid stage of LikeliOptimizer.find on the GPU(s).  With torch.distributed initialised the
        candidate range is sharded contiguously over the ranks and only the (nll, index) pairs cross
        NCCL; every rank returns the same (best_param, best_idx[, local nll])."""
        T, P = int(theta_sections), int(p_sections)
        total = (T ** self._k) * (P ** self._k)
        seeded = getattr(self, '_grid_guess', None)
        if seeded is not None and (T, P) in seeded and not return_nll:   # set by a leave-one-out batch (loo_grid_guesses)
            return seeded[(T, P)]
        rank, world = dist.rank_world(distributed)
        lo, hi = dist.shard_range(total, rank, world)
        bv_d, bi_d, nll = self.likelihood_grid(T, P, lo, hi - lo, return_nll=return_nll, as_device=True)
        bv, bi = dist.global_first_min_dev(bv_d, bi_d)  # one 16-byte all_gather + one 16-byte read-back
        param = None if bi < 0 else grid_param(self._k, T, P, bi)
        if return_nll:
            return param, bi, nll.cpu().numpy()
        return param, bi

    def loo_grid_guesses(self, theta_sections=8, p_sections=5, return_nll=False):
        """Grid stage of re-training one model per left-out sample (PRESS, validation.py:79-110) as ONE batched launch
        over (sample set, candidate) pairs -- sm_krig_loo_grid.  Returns a list of n (best_param, best_idx) pairs: entry m
        is what Kriging(X without sample m, y without sample m).grid_search(...) returns."""
        L = _cabi.lib()
        d = self._dev()
        T, P = int(theta_sections), int(p_sections)
        total = (T ** self._k) * (P ** self._k)
        th_axis, p_axis = grid_axes(T, P)
        th_lin = device.to_dev(np.array([10.**e for e in th_axis]))
        pa = device.to_dev(p_axis)
        nll = device.empty((self._n, total)) if return_nll else None
        bv, bi = device.empty(self._n), device.empty(self._n, dtype=_int64())
        _cabi.check(L.sm_krig_loo_grid(device.ptr(d['X']), device.ptr(d['y']), self._n, self._k, device.ptr(th_lin), T,
                                       device.ptr(pa), P, device.ptr(nll), device.ptr(bv), device.ptr(bi), device.stream()))
        idx = bi.cpu().numpy()
        out = [(None if c < 0 else grid_param(self._k, T, P, int(c)), int(c)) for c in idx]
        if return_nll:
            return out, nll.cpu().numpy()
        return out

    def predict_batch(self, x_query, return_mse=False, distributed=None, gather=True):
        """Mean (and ordinary-Kriging MSE) at many query points: Xq[Nq,k] -> y[Nq] (, mse[Nq]).
        Replaces a Python loop over Kriging.predict (kriging.py:206-220).
        With torch.distributed initialised (and distributed is not False) every rank, holding the same fitted model,
        predicts the contiguous row range [lo, hi) of Xq; gather=True all-gathers the slices so every rank returns all
        Nq results, gather=False returns this rank's rows only (results stay sharded)."""
        if self._fit is None and not self._fit_stale:  # the reference predicts with whatever update_param left behind
            raise _cabi.SurrogateB200Error('predict before update_param / train: no fitted state')
        self._ensure_fit()
        xq = np.asarray(x_query, dtype=np.float64)
        if xq.ndim == 1:
            xq = xq.reshape(-1, self._k)
        nq = xq.shape[0]
        rank, world = dist.rank_world(distributed)
        lo, hi = dist.shard_range(nq, rank, world)
        y, mse = self.predict_device(device.to_dev(xq[lo:hi]), return_mse)
        if world > 1 and gather:
            y = dist.gather_slices(y, nq)
            mse = dist.gather_slices(mse, nq) if return_mse else None
        if return_mse:
            return y.cpu().numpy(), mse.cpu().numpy()
        return y.cpu().numpy()

    def predict_mesh(self, lo, hi, m, begin=0, count=None, return_mse=True, as_device=False):
        """Value (and MSE) at points [begin, begin + count) of the full-factorial mesh with m[d] points per dimension over
        [lo[d], hi[d]] (last dimension fastest) -- the surface-plot meshes of the reference (examples/kriging_r3.py:59-70,
        PlotHelper.plot_function_3d) and BASELINE config 5 -- without materialising the points: sm_krig_predict_mesh
        tabulates the per-dimension correlation factors over the axes.  Within k ulp of predict_batch on device.mesh_points(...)."""
        L = _cabi.lib()
        if return_mse:
            self._ensure_mse_state()
        else:
            self._ensure_fit()
        d, f = self._dev(), self._fit
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        m = np.ascontiguousarray(m, dtype=np.int64)
        count = int(np.prod(m)) - int(begin) if count is None else int(count)
        y = device.empty(count)
        mse = device.empty(count) if return_mse else None
        if not f['ok']:
            y.fill_(float('nan'))
            if mse is not None:
                mse.fill_(float('nan'))
        else:
            want_mse = return_mse and f.get('info', 0) != 3
            _cabi.check(L.sm_krig_predict_mesh(device.ptr(d['X']), self._n, self._k, device.ptr(f['theta']), device.ptr(f['p']),
                                               device.ptr(f['w']), f['mu'], f['sigma2'], device.ptr(f['Mop']) if want_mse else None,
                                               device.ptr(f['b']), device.ptr(f['rres']), lo.ctypes.data, hi.ctypes.data,
                                               m.ctypes.data, int(begin), count, device.ptr(y),
                                               device.ptr(mse) if want_mse else None, device.stream()))
            if return_mse and not want_mse:
                mse.fill_(float('nan'))
        if as_device:
            return (y, mse) if return_mse else y
        return (y.cpu().numpy(), mse.cpu().numpy()) if return_mse else y.cpu().numpy()

    def predict_device(self, xq_dev, return_mse=False):
        """predict_batch on a device tensor Xq[Nq,k]; results stay on the device (no sync)."""
        L = _cabi.lib()
        if return_mse:
            self._ensure_mse_state()
        else:
            self._ensure_fit()
        d, f = self._dev(), self._fit
        nq = xq_dev.shape[0]
        y = device.empty(nq)
        mse = device.empty(nq) if return_mse else None
        if not f['ok']:
            y.fill_(float('nan'))
            if mse is not None:
                mse.fill_(float('nan'))
            return y, mse
        want_mse = mse is not None and f.get('info', 0) != 3
        _cabi.check(L.sm_krig_predict(device.ptr(d['X']), self._n, self._k, device.ptr(f['theta']), device.ptr(f['p']),
                                      device.ptr(f['w']), f['mu'], f['sigma2'], device.ptr(f['Mop']) if want_mse else None,
                                      device.ptr(f['b']), device.ptr(f['rres']), device.ptr(xq_dev), nq,
                                      device.ptr(y), device.ptr(mse) if want_mse else None, device.stream()))
        if mse is not None and not want_mse:
            mse.fill_(float('nan'))  # info 3: weights from the double-double factorisation, but no fp64 factor for L^-1
        return y, mse

    # ------------------------------------------------------------------ likelihood surface plots
    def plot_theta_likelihood_r2(self, ax=None, pgf=False, opti_path=[]):
        """colormap of the likelihood over theta1, theta2 (reference: kriging.py:222-252); the 50 x 50
        sweep is one batched launch instead of 2500 update_param calls."""
        if self._k != 2:
            print('ERROR: plot_theta_likelihood_R2 only works with exactly 2 inputs')
            return
        from myutils.plot_helper import PlotHelper
        thetas = np.logspace(-5, 5, num=50)
        e1, e2 = np.meshgrid(np.log10(thetas), np.log10(thetas), indexing='ij')
        cand_e = np.stack([e1.ravel(), e2.ravel()], axis=1)
        cand_p = np.tile(self._p, (cand_e.shape[0], 1))
        likely_thet = self.likelihood_batch(cand_e, cand_p).reshape(50, 50).T
        plt_theta = PlotHelper([r'$\theta_{1}$', r'$\theta_{2}$'], fancy=False, ax=ax, pgf=pgf)
        plt_theta.ax.set_xscale('log')
        plt_theta.ax.set_yscale('log')
        pcol = plt_theta.ax.pcolor(thetas, thetas, likely_thet, cmap='YlOrRd_r')
        if len(opti_path) > 0:
            plt_theta.ax.plot(10**opti_path[:, 0], 10**opti_path[:, 1], '+', color='white', markeredgewidth=0.5,
                              markersize=5, label='Optimierer-Pfad')
        plt_theta.ax.plot(self._theta[0], self._theta[1], 'x', color='black', label='Minimum', markersize=8,
                          markeredgewidth=1.5)
        plt_theta.finalize(width=6, height=5, legendLoc=4, show_legend=False)
        return pcol

    def plot_p_likelihood_r2(self, ax=None, pgf=False, opti_path=[]):
        """colormap of the likelihood over p1, p2 (reference: kriging.py:254-282)"""
        if self._k != 2:
            print('ERROR: plot_p_likelihood_R2 only works with exactly 2 inputs')
            return
        from myutils.plot_helper import PlotHelper
        ps = np.linspace(1., 2., num=50)
        p1, p2 = np.meshgrid(ps, ps, indexing='ij')
        cand_p = np.stack([p1.ravel(), p2.ravel()], axis=1)
        cand_e = np.tile(np.log10(self._theta), (cand_p.shape[0], 1))
        likely_p = self.likelihood_batch(cand_e, cand_p).reshape(50, 50).T
        plt_P = PlotHelper(['$p_{1}$', '$p_{2}$'], fancy=False, ax=ax, pgf=pgf)
        pcol = plt_P.ax.pcolor(ps, ps, likely_p, cmap='YlOrRd_r')
        if len(opti_path) > 0:
            plt_P.ax.plot(opti_path[:, 0], opti_path[:, 1], '+', color='white', markeredgewidth=0.5, markersize=5,
                          label='Optimierer-Pfad')
        plt_P.ax.plot(self._p[0], self._p[1], 'x', color='k', label='Minimum', markersize=8, markeredgewidth=1.5)
        plt_P.finalize(width=6, height=5, legendLoc=4, show_legend=False)
        return pcol

    def plot_likelihoods(self, fancy=False, pgf=False, opti_path=[]):
        """both colormaps side by side (reference: kriging.py:284-327)"""
        if self._k != 2:
            print('ERROR: plot_p_likelihood_R2 only works with exactly 2 inputs')
            return
        from myutils.plot_helper import PlotHelper
        path_p, path_theta = [], []
        if len(opti_path) > 0:
            path_theta = np.array([opti_path[:, 0], opti_path[:, 1]]).T
            path_p = np.array([opti_path[:, 2], opti_path[:, 3]]).T
        plt_like = PlotHelper([], fancy=fancy, pgf=pgf)
        import matplotlib.pyplot as plt
        plt_like.fig = plt.figure()
        ax1 = plt_like.fig.add_subplot(121)
        ax2 = plt_like.fig.add_subplot(122)
        pcol1 = self.plot_p_likelihood_r2(ax=ax1, pgf=pgf, opti_path=path_p)
        pcol2 = self.plot_theta_likelihood_r2(ax=ax2, pgf=pgf, opti_path=path_theta)
        like_min = min(pcol1.get_array().min(), pcol2.get_array().min())
        like_max = max(pcol1.get_array().max(), pcol2.get_array().max())
        pcol1.set_clim(like_min, like_max)
        pcol2.set_clim(like_min, like_max)
        plt_like.fig.set_size_inches(6, 3)
        plt_like.fig.colorbar(pcol2)
        return plt_like

    def get_p(self):
        return self._p

    def get_theta(self):
        return self._theta


def _int64():
    import torch
    return torch.int64


def grid_axes(theta_sections, p_sections):
    """axes of LikeliOptimizer.generate_grid (likeli_optimizer.py:53-55): theta axis holds EXPONENTS"""
    return np.linspace(-5, 5, num=theta_sections), np.linspace(1., 2., num=p_sections)


def grid_param(k, theta_sections, p_sections, c):
    """parameter vector [log10 theta_1..k, p_1..k] of flat grid index c (likeli_optimizer.py:59-86:
    c = ip * T^k + it, dimension 0 the fastest digit)."""
    th_axis, p_axis = grid_axes(theta_sections, p_sections)
    T, P = theta_sections, p_sections
    ip, it = divmod(int(c), T ** k)
    param = [th_axis[(it // T ** d) % T] for d in range(k)]
    param += [p_axis[(ip // P ** d) % P] for d in range(k)]
    return param


# ---- helper objects for scipy.optimize.basinhopping (reference: kriging.py:336-365) -----------------------------------


class BasinHoppingBounds(object):
    """accept_test: a hop is kept only if every coordinate lies inside [xmin, xmax]"""

    def __init__(self, xmax=[5., 5., 2., 2.], xmin=[-5., -5., 1., 1.]):
        self.xmax = np.array(xmax)
        self.xmin = np.array(xmin)

    def __call__(self, **kwargs):
        x = kwargs["x_new"]
        return bool(np.all(x <= self.xmax) and np.all(x >= self.xmin))


class BasinHoppingStep(object):
    """take_step: a completely fresh random point.  Quirk of the reference kept on purpose (kriging.py:357-365,
    SURVEY.md App. A.8): the first half of x is drawn as 10**U(-5, 5) although the objective reads those slots as
    EXPONENTS, so BasinHoppingBounds throws most proposals away; the second half is p ~ U(1, 2)."""

    def __init__(self, stepsize=1.):
        self.stepsize = stepsize

    def __call__(self, x):
        half = len(x) / 2
        for slot in range(len(x)):
            x[slot] = 10 ** np.random.uniform(-5, 5) if slot < half else np.random.uniform(1., 2)
        return x
