"""Model-independent validation measures -- drop-in for the reference's ``mylibs/validation.py``.

Host-side caller of the surrogate protocol (``Cls(x, fx)``, ``update_param``, ``train``, ``predict``); it is
the code that *defines* the drop-in boundary (reference: validation.py:96-108).  Same class, method names,
argument order and return values.  When ``surro_func`` is the bound ``predict`` of a B200 surrogate the
per-point Python loops of the reference become one ``predict_batch`` launch; any other callable is mapped
point by point like the reference does.
"""
import math

import numpy as np


def _predict_all(surro_func, points):
    """predictions for a list of points: batched when the surrogate offers it, else one call per point"""
    owner = getattr(surro_func, '__self__', None)
    if owner is not None and getattr(surro_func, '__name__', '') == 'predict' and hasattr(owner, 'predict_batch'):
        pts = np.asarray(points, dtype=np.float64)
        if pts.ndim == 1:
            pts = pts.reshape(-1, 1)
        return np.asarray(owner.predict_batch(pts), dtype=np.float64)
    return np.array([surro_func(pt) for pt in points], dtype=np.float64)


class Validation:

    def __init__(self):
        pass

    def calc_deviation(self, params, values, surro_func):
        """mean absolute deviation relative to the mean of `values` (reference: validation.py:23-40)"""
        values = np.asarray(values, dtype=np.float64)
        pred = _predict_all(surro_func, params)
        return np.mean(np.abs(values - pred)) / values.mean()

    def calc_rmse(self, vali_x, vali_fx, surro_func):
        """root mean square error (reference: validation.py:42-53)"""
        err = np.asarray(vali_fx, dtype=np.float64) - _predict_all(surro_func, vali_x)
        return math.sqrt(float(np.sum(err ** 2)) / len(vali_x))

    def calc_mae(self, vali_x, vali_fx, surro_func):
        """maximum absolute error (reference: validation.py:55-67); -inf for an empty set, like the reference"""
        if len(vali_x) == 0:
            return -1 * float('inf')
        return float(np.max(np.abs(np.asarray(vali_fx, dtype=np.float64) - _predict_all(surro_func, vali_x))))

    def calc_rae(self, vali_x, vali_fx, surro_func):
        """relative absolute error per validation point (reference: validation.py:69-77)"""
        vali_fx = np.asarray(vali_fx, dtype=np.float64)
        return abs(np.divide((vali_fx - _predict_all(surro_func, vali_x)), vali_fx))

    def calc_press(self, known_x, known_fx, surro_func, surro_class, update_params=None, quirk=True):
        """prediction sum of squares (reference: validation.py:79-110).

        quirk=True (default, the reference's behaviour): the reference re-trains a surrogate for every i but ALWAYS leaves
        out sample 0 (np.delete(x, 0, axis=0), validation.py:92,94), i.e. it trains the same model n times.  That model
        is built once here and evaluated at all n points in one batch -- the returned value is the same.
        quirk=False: true leave-one-out -- model i is trained without sample i and evaluated at sample i.  For the B200
        Kriging class the grid stage of all n trainings is ONE batched launch (Kriging.loo_grid_guesses, n sample sets x
        1600 candidates); the SLSQP refinements and the n predictions follow per model.
        Any exception while building a model gives the reference's warning and 0."""
        n = len(known_x)
        res = _predict_all(surro_func, known_x)
        x_all, fx_all = np.array(known_x), np.array(known_fx)

        def build(x, fx, seed_guess=None):
            sur = surro_class(x, fx)
            if update_params is not None:
                if len(update_params) == 1:
                    sur.update_param(update_params[0])
                elif len(update_params) == 2:
                    sur.update_param(update_params[0], update_params[1])
                else:
                    print('Validation.calc_press please update to support more params than 2')
            if seed_guess is not None:
                sur._grid_guess = {(8, 5): seed_guess}   # the grid of LikeliOptimizer.find (likeli_optimizer.py:36)
            sur.train()
            return sur

        try:
            if quirk:
                sur = build(np.delete(x_all, 0, axis=0), np.delete(fx_all, 0, axis=0))
                res2 = _predict_all(sur.predict, known_x)
            else:
                guesses = [None] * n
                if hasattr(surro_class, 'loo_grid_guesses'):
                    guesses = surro_class(x_all, fx_all).loo_grid_guesses(8, 5)
                res2 = np.empty(n)
                for i in range(n):
                    sur = build(np.delete(x_all, i, axis=0), np.delete(fx_all, i, axis=0), guesses[i])
                    res2[i] = sur.predict(known_x[i])
        except Exception as e:
            print('WARNING: missing train() method: ' + str(e))
            return 0.
        return math.sqrt(float(np.sum((res - res2) ** 2.)) / n)

    def run_full_analysis(self, params, values, known_x, known_fx, vali_x, vali_fx, surro_func, surro_class,
                          update_params=None):
        """all of the above (reference: validation.py:112-132)"""
        res = ValidationResults()
        res.deviation = self.calc_deviation(params, values, surro_func)
        res.rmse = self.calc_rmse(vali_x, vali_fx, surro_func)
        res.mae = self.calc_mae(vali_x, vali_fx, surro_func)
        res.rae = self.calc_rae(vali_x, vali_fx, surro_func)
        res.press = self.calc_press(known_x, known_fx, surro_func, surro_class, update_params=update_params)
        return res


class ValidationResults():
    def __init__(self):
        self.deviation = 0.
        self.rmse = 0.
        self.mae = 0.
        self.rae = 0.
        self.press = 0.
