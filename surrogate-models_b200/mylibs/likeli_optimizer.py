"""Grid + SLSQP search for the minimum negative log-likelihood -- drop-in for the reference's
``mylibs/likeli_optimizer.py`` (same class, attributes, method names and return values).

``generate_grid`` keeps the reference's enumeration (flat index c = ip * T^k + it, dimension 0 the fastest digit,
first strict minimum wins, NaN never wins; reference :50-97).  When ``func`` is the bound
``_calc_likelihood_opti_exp`` of a B200 ``Kriging`` the whole grid is ONE batched kernel launch
(``Kriging.grid_search``); any other callable is evaluated candidate by candidate in the same order.

Additive (not in the reference): ``multi_start`` -- SLSQP is started from the best ``multi_start`` grid candidates
instead of only the first minimum and the best end point is returned.  The default 1
is the reference's behaviour.
"""
from scipy.optimize import minimize
import numpy as np

THETA_EXP_RANGE = (-5, 5)   # grid axis and SLSQP bounds of log10(theta) (reference :39,:55)
P_RANGE = (1., 2.)          # grid axis and SLSQP bounds of p (reference :41,:53)


class LikeliOptimizer:

    def __init__(self, debug=False):
        """(reference: likeli_optimizer.py:21-25; pBounds / thetaBoundsExp are unused there as well)"""
        self.debug = debug
        self.pBounds = (1., 2.)
        self.thetaBoundsExp = (-5, 3)
        self.maxIter = 5e3
        self.multi_start = 1

    # -------------------------------------------------------------------------------------------- SLSQP stage
    def find(self, func, dimensions):
        """
        finds the min neg. log. likelihood: grid guess (8 theta sections x 5 p sections per dimension), then
        scipy SLSQP from it inside [-5, 5]^k x [1, 2]^k (reference: likeli_optimizer.py:27-48)
        :param func: pointer to the likelihood calculation function
        :param dimensions: number of dimensions
        :return: the minimum as a scipy.optimize.minimize result
        """
        box = [THETA_EXP_RANGE] * dimensions + [P_RANGE] * dimensions
        settings = dict(disp=False, maxiter=self.maxIter)
        owner = self._batched_owner(func)
        batched = owner is not None and dimensions == owner._k
        if batched and getattr(owner, 'batched_stencil', False):
            # SciPy's own forward-difference stencil (same points, same steps), evaluated as ONE batched launch
            settings['workers'] = owner._objective_map
        starts = [self.generate_grid(func, dimensions, 8, 5)]
        if batched and self.multi_start > 1:
            starts = owner.grid_top(8, 5, self.multi_start) or starts
        best = None
        for x0 in starts:
            res = minimize(func, x0, method='SLSQP', tol=1e-8, options=settings, bounds=box)
            if res.nit >= self.maxIter:
                print('WARNING: max iter was used (LikeliOptimizer.py)')
            if best is None or res.fun < best.fun:
                best = res
        return best

    @staticmethod
    def _batched_owner(func):
        """the Kriging instance behind `func` if the batched GPU grid can stand in for the loop"""
        owner = getattr(func, '__self__', None)
        if owner is not None and getattr(func, '__name__', '') == '_calc_likelihood_opti_exp' \
                and hasattr(owner, 'grid_search'):
            return owner
        return None

    # -------------------------------------------------------------------------------------------- grid stage
    def generate_grid(self, func, dimensions, theta_sections, p_sections):
        """(reference: likeli_optimizer.py:50-87) returns the parameter list of the first strict minimum,
        or None when every candidate is inf / NaN."""
        from mylibs.kriging import grid_param
        total = (theta_sections ** dimensions) * (p_sections ** dimensions)
        owner = self._batched_owner(func)
        if owner is not None and dimensions == owner._k:
            records = getattr(owner, 'records', None)
            param, _ = owner.grid_search(theta_sections, p_sections)
            if records is not None:  # evaluation order of the reference: every grid row, in order
                records.extend(np.array(grid_param(dimensions, theta_sections, p_sections, c)) for c in range(total))
            return param
        # opaque callable: one call per candidate, walking the odometers exactly as the reference does so that
        # `increase_i` keeps its role (an override of it changes the walk here as it would there)
        winner, winner_val = None, float('inf')
        p_digits = np.zeros(dimensions)
        for _ in range(p_sections ** dimensions):
            t_digits = np.zeros(dimensions)
            for _ in range(theta_sections ** dimensions):
                c = sum(int(t_digits[d]) * theta_sections ** d for d in range(dimensions)) \
                    + (theta_sections ** dimensions) * sum(int(p_digits[d]) * p_sections ** d for d in range(dimensions))
                candidate = grid_param(dimensions, theta_sections, p_sections, c)
                value = func(candidate)
                if value < winner_val:  # strict: the first minimum stays, NaN never compares true
                    winner, winner_val = candidate, value
                t_digits = self.increase_i(t_digits, theta_sections)
            p_digits = self.increase_i(p_digits, p_sections)
        return winner

    def increase_i(self, i_list, max_len):
        """advance a little-endian odometer with digits 0..max_len-1 by one, in place; wraps to all zeros
        (reference: likeli_optimizer.py:89-97)"""
        pos = 0
        while pos < len(i_list) and i_list[pos] >= max_len - 1:
            i_list[pos] = 0
            pos += 1
        if pos < len(i_list):
            i_list[pos] += 1
        return i_list
