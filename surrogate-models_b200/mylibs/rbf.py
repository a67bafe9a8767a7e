"""n-dimensional radial basis function interpolation on one B200 -- drop-in for the reference's
``mylibs/rbf.py``.  Same class / methods / kernel names; the weight solve and the evaluation run in
sm_100a kernels (``sm_rbf_fit`` / ``sm_rbf_predict``).  Additive: ``predict_batch``.
"""
import math

import numpy as np

from sm_b200 import _cabi, device


class RBF:

    def __init__(self, known_in, known_val):
        """
        :param known_in: list of lists with input sample points
        :param known_val: list of results for the known_in
        (reference: rbf.py:21-34)
        """
        self._known_in = np.array(known_in, dtype=np.float64)
        self._known_val = np.array(known_val, dtype=np.float64)
        if len(self._known_in.shape) == 1:
            self._known_in = self._known_in.reshape((self._known_in.shape[0], 1))
        self._k = self._known_in.shape[1]
        self._n = self._known_in.shape[0]
        self._coeff = None
        self._rbf_const = 1.
        self._rbf = gaus_rbf
        self._d = None

    def _dev(self):
        if self._d is None:
            device.require_cuda()
            self._d = dict(X=device.to_dev(self._known_in), y=device.to_dev(self._known_val), coeff=None)
        return self._d

    def train(self):
        """no-op, like the reference (rbf.py:36-41): coefficients only exist after update_param"""
        pass

    def update_param(self, rbf_const, rbf_name):
        """
        updates the parameters of the surrogate model (reference: rbf.py:43-65)
        :param rbf_const: the rbf constant a
        :param rbf_name: the short name of the rbf to use (lin, cubic, gaus, mq, imq)
        """
        self._rbf_const = rbf_const
        if rbf_name == 'lin':
            self._rbf = lin_rbf
        elif rbf_name == 'cubic':
            self._rbf = cubic_rbf
        elif rbf_name == 'gaus':
            self._rbf = gaus_rbf
        elif rbf_name == 'multi-quadratic' or rbf_name == 'mq':
            self._rbf = multi_quad_rbf
        elif rbf_name == 'inverse-multi-quadratic' or rbf_name == 'imq':
            self._rbf = inv_multi_quad_rbf
        else:
            print('WARNING: unknown rbf_name (' + rbf_name + '), I will just use gaus for you.')
            print('next time chose one of ["gaus", "multi-quadratic", "inverse-multi-quadratic"]')
            self._rbf = gaus_rbf
        self._calc_coefficiants()

    def _kind(self):
        return _KIND[self._rbf]

    def _calc_coefficiants(self):
        """coeff = Phi^-1 y on the device (reference: rbf.py:67-80).  A singular Phi raises
        numpy.linalg.LinAlgError, as np.linalg.solve does there."""
        import torch
        L = _cabi.lib()
        d = self._dev()
        coeff = device.empty(self._n)
        info = torch.zeros(1, dtype=torch.int32, device=coeff.device)
        _cabi.check(L.sm_rbf_fit(device.ptr(d['X']), device.ptr(d['y']), self._n, self._k, self._kind(),
                                 float(self._rbf_const), device.ptr(coeff), device.ptr(info), device.stream()))
        if int(info.item()) > 0:
            raise np.linalg.LinAlgError('Singular matrix')
        d['coeff'] = coeff
        self._coeff = coeff.cpu().numpy()
        return self._coeff

    def get_coeff(self):
        return self._coeff

    def predict(self, x_pred):
        """
        predicts a value from the surrogate model (reference: rbf.py:85-103)
        :param x_pred: vector of input values
        :return: result value (Python float)
        """
        xq = np.array([[x_pred[i] for i in range(self._k)]], dtype=np.float64)
        return self.predict_batch(xq)[0].item()

    def predict_batch(self, x_query, distributed=None, gather=True):
        """f(x) for many query points: Xq[Nq,k] -> y[Nq] (replaces a loop over predict).  With torch.distributed
        initialised the query rows are sharded contiguously over the ranks (each holds the same coefficients);
        gather=True returns all Nq values on every rank, gather=False this rank's rows only."""
        from sm_b200 import dist
        xq = np.asarray(x_query, dtype=np.float64)
        if xq.ndim == 1:
            xq = xq.reshape(-1, self._k)
        nq = xq.shape[0]
        rank, world = dist.rank_world(distributed)
        lo, hi = dist.shard_range(nq, rank, world)
        y = self.predict_device(device.to_dev(xq[lo:hi]))
        if world > 1 and gather:
            y = dist.gather_slices(y, nq)
        return y.cpu().numpy()

    # ------------------------------------------------------------------ batch over the rbf constant (additive)
    def sweep_rbf_const(self, a_values, x_query, rbf_name='gaus'):
        """Predictions of the surrogate at x_query for EVERY rbf constant in a_values: one weight solve + one evaluation
        per constant, enqueued back to back on the stream with a single synchronisation at the end -- the batched form of
        the loop `update_param(a, name); predict(x)` that Surrogate.auto_fit_rbf drives through SLSQP
        (wingconstruction/surrogate_run.py:146-166).  Returns pred[len(a_values), Nq]; a singular system gives a NaN row.
        The object's own parameters are left untouched."""
        import torch
        L = _cabi.lib()
        d = self._dev()
        kind = _KIND[_NAMES.get(rbf_name, gaus_rbf)]
        xq = np.asarray(x_query, dtype=np.float64)
        if xq.ndim == 1:
            xq = xq.reshape(-1, self._k)
        xq_d = device.to_dev(xq)
        a_values = [float(a) for a in a_values]
        pred = device.empty((len(a_values), xq.shape[0]))
        coeff = device.empty((len(a_values), self._n))
        info = torch.zeros(len(a_values), dtype=torch.int32, device=pred.device)
        for i, a in enumerate(a_values):
            _cabi.check(L.sm_rbf_fit(device.ptr(d['X']), device.ptr(d['y']), self._n, self._k, kind, a,
                                     coeff[i].data_ptr(), info[i:].data_ptr(), device.stream()))
            _cabi.check(L.sm_rbf_predict(device.ptr(d['X']), self._n, self._k, kind, a, coeff[i].data_ptr(),
                                         device.ptr(xq_d), xq.shape[0], pred[i].data_ptr(), device.stream()))
        out = pred.cpu().numpy()
        out[info.cpu().numpy() > 0] = np.nan
        return out

    def auto_fit(self, vali_x, vali_fx, rbf_name='gaus', init_a=1., bounds=(0.01, 5.), grid=48, scale=1.):
        """The search of Surrogate.auto_fit_rbf (surrogate_run.py:146-166: SLSQP over the rbf constant a in [0.01, 5],
        objective = validation RMSE / scale) with the B200 batch in front: the RMSE over a `grid`-point sweep of a comes
        from ONE sweep_rbf_const call, SLSQP (same settings as the reference) then starts from the better of the sweep's
        minimum and init_a.  Leaves the model trained at the optimum and returns (a, rmse / scale)."""
        from scipy.optimize import minimize
        vali_x = np.asarray(vali_x, dtype=np.float64)
        vali_fx = np.asarray(vali_fx, dtype=np.float64)

        def objective(a):
            self.update_param(float(np.atleast_1d(a)[0]), rbf_name)
            err = vali_fx - self.predict_batch(vali_x)
            return math.sqrt(float(np.sum(err ** 2)) / len(vali_x)) / scale
        cand = np.concatenate([[init_a], np.linspace(bounds[0], bounds[1], grid)])
        sweep = self.sweep_rbf_const(cand, vali_x, rbf_name)
        rmse = np.sqrt(np.sum((vali_fx[None, :] - sweep) ** 2, axis=1) / len(vali_x)) / scale
        start = float(cand[int(np.nanargmin(rmse))])
        res = minimize(objective, start, method='SLSQP', tol=1e-8, options={'disp': False, 'maxiter': 999}, bounds=[bounds])
        best = float(res.x[0])
        if not (res.fun <= np.nanmin(rmse)):  # SLSQP may end worse than its start (the reference keeps res.x regardless)
            best = start
        self.update_param(best, rbf_name)
        return best, objective(best)

    def predict_device(self, xq_dev):
        if self._coeff is None:
            raise TypeError("'NoneType' object is not subscriptable (predict before update_param)")
        L = _cabi.lib()
        d = self._dev()
        if d['coeff'] is None:
            d['coeff'] = device.to_dev(self._coeff)
        nq = xq_dev.shape[0]
        y = device.empty(nq)
        _cabi.check(L.sm_rbf_predict(device.ptr(d['X']), self._n, self._k, self._kind(), float(self._rbf_const),
                                     device.ptr(d['coeff']), device.ptr(xq_dev), nq, device.ptr(y), device.stream()))
        return y


# scalar kernel definitions kept for API compatibility (reference: rbf.py:105-118); the device kernels
# implement the same formulas
def lin_rbf(a, r):
    return r


def cubic_rbf(a, r):
    return r**3


def gaus_rbf(a, r):
    return math.e**(-((a*r)**2))


def multi_quad_rbf(a, r):
    return math.sqrt(1 + (a * r) ** 2)


def inv_multi_quad_rbf(a, r):
    return (1+r**2)**(a/2)


_KIND = {lin_rbf: 0, cubic_rbf: 1, gaus_rbf: 2, multi_quad_rbf: 3, inv_multi_quad_rbf: 4}
_NAMES = {'lin': lin_rbf, 'cubic': cubic_rbf, 'gaus': gaus_rbf, 'multi-quadratic': multi_quad_rbf, 'mq': multi_quad_rbf,
          'inverse-multi-quadratic': inv_multi_quad_rbf, 'imq': inv_multi_quad_rbf}
