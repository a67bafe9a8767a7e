"""Grid + SLSQP search for the minimum negative log-likelihood -- drop-in for the reference's
``mylibs/likeli_optimizer.py``.

``generate_grid`` keeps its signature and enumeration (outer p-odometer, inner theta-odometer, dimension
0 the fastest digit, first strict minimum wins; reference :50-97).  When ``func`` is the bound
``_calc_likelihood_opti_exp`` of a B200 ``Kriging`` the whole grid is ONE batched kernel launch
(``Kriging.grid_search``); for any other callable it loops exactly like the reference does.
"""
from scipy.optimize import minimize
import numpy as np


class LikeliOptimizer:

    def __init__(self, debug=False):
        """(reference: likeli_optimizer.py:21-25)"""
        self.debug = debug
        self.pBounds = (1., 2.)
        self.thetaBoundsExp = (-5, 3)
        self.maxIter = 5e3

    def find(self, func, dimensions):
        """
        finds the min neg. log. likelihood by generating a grid and starting gradient based optimization
        from its best result (reference: likeli_optimizer.py:27-48)
        :param func: pointer to the likelihood calculation function
        :param dimensions: number of dimensions
        :return: the minimum as a scipy.optimize.minimize result
        """
        opt = {'disp': False, 'maxiter': self.maxIter}
        guess = self.generate_grid(func, dimensions, 8, 5)
        bnds = []
        for i in range(0, dimensions):
            bnds.append((-5, +5))
        for i in range(0, dimensions):
            bnds.append((1., 2.))
        owner = self._batched_owner(func)
        if owner is not None and dimensions == owner._k and getattr(owner, 'batched_stencil', False):
            # SciPy's own forward-difference stencil (same points, same steps), evaluated as ONE batched launch
            opt['workers'] = owner._objective_map
        minima_res = minimize(func, guess, method='SLSQP', tol=1e-8, options=opt, bounds=bnds)
        if minima_res.nit >= self.maxIter:
            print('WARNING: max iter was used (LikeliOptimizer.py)')
        return minima_res

    @staticmethod
    def _batched_owner(func):
        """the Kriging instance behind `func` if the batched GPU grid can stand in for the loop"""
        owner = getattr(func, '__self__', None)
        if owner is not None and getattr(func, '__name__', '') == '_calc_likelihood_opti_exp' \
                and hasattr(owner, 'grid_search'):
            return owner
        return None

    def generate_grid(self, func, dimensions, theta_sections, p_sections):
        """(reference: likeli_optimizer.py:50-87) returns the parameter list of the first strict minimum,
        or None when every candidate is inf / NaN."""
        owner = self._batched_owner(func)
        if owner is not None and dimensions == owner._k:
            records = getattr(owner, 'records', None)
            param, best_idx = owner.grid_search(theta_sections, p_sections)
            if records is not None:  # evaluation order of the reference: every grid row, in order
                from mylibs.kriging import grid_param
                total = (theta_sections ** dimensions) * (p_sections ** dimensions)
                records.extend(np.array(grid_param(dimensions, theta_sections, p_sections, c)) for c in range(total))
            return param
        minima_val = float('inf')
        minima_param = None
        ps = np.linspace(1., 2., num=p_sections)
        thetas = np.linspace(-5, 5, num=theta_sections)
        p_param_i = np.zeros((dimensions))
        for ip in range(len(ps)**dimensions):
            t_param_i = np.zeros((dimensions))
            for it in range(len(thetas)**dimensions):
                param = []
                for i in range(0, dimensions):
                    param.append(thetas[int(t_param_i[i])])
                for i in range(0, dimensions):
                    param.append(ps[int(p_param_i[i])])
                res_val = func(param)
                if res_val < minima_val:
                    minima_val = res_val
                    minima_param = param
                t_param_i = self.increase_i(t_param_i, len(thetas))
            p_param_i = self.increase_i(p_param_i, len(ps))
        return minima_param

    def increase_i(self, i_list, max_len):
        """little-endian odometer increment (reference: likeli_optimizer.py:89-97)"""
        for i in range(0, len(i_list)):
            if i_list[i] >= max_len - 1:
                i_list[i] = 0
            else:
                i_list[i] += 1
                return i_list
        return i_list
