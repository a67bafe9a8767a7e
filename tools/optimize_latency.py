"""optimize() wall time (grid + SLSQP) for small models, under the current SMB_LAZY_FIT_MIN_N"""
import io, os, sys, time, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import numpy as np, torch
from mylibs.kriging import Kriging, LAZY_FIT_MIN_N
from sm_b200 import workloads, _cabi
cases = {'c1 n=8 k=1': workloads.c1_kriging_r2(), 'c5 n=36 k=2': workloads.c5_kriging_r3(), 'c2 n=64 k=2': workloads.c2_halton64('fb')}
rng = np.random.default_rng(0); X = rng.uniform(0, 1, (100, 4)); cases['n=100 k=4'] = (X, np.sin(3 * X.sum(axis=1)))
for name, (X, y) in cases.items():
    k = Kriging(X, y)
    with contextlib.redirect_stdout(io.StringIO()):
        k.optimize()
    ts = []
    for _ in range(3):
        k = Kriging(X, y); n0 = _cabi.launch_count(); t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            k.optimize(record_data=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f'LAZY_FIT_MIN_N={LAZY_FIT_MIN_N} {name}: optimize {min(ts)*1e3:.1f} ms, {len(k.records)} evaluations, {_cabi.launch_count()-n0} launches, nll {k.calc_likelihood():.9f} theta {k._theta} p {k._p}')
