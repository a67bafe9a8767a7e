"""Small invocations of every kernel added in round 2, for compute-sanitizer (memcheck / racecheck / synccheck):
compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import numpy as np, torch
from mylibs.kriging import Kriging
from mylibs.rbf import RBF
from sm_b200 import workloads, device, _cabi
rng = np.random.default_rng(0)
X5, y5 = workloads.c5_kriging_r3()
k = Kriging(X5, y5)
print('grid n=36', k.grid_search(8, 5))                      # regs NT=5 + double-double pass (shared memory)
print('loo', k.loo_grid_guesses(4, 2)[:2])                   # model dimension
for n, kk in ((64, 2), (57, 3), (80, 2), (73, 4), (96, 2), (150, 2), (200, 3)):
    X = rng.uniform(0, 10, (n, kk)); y = np.sin(X.sum(axis=1))
    kr = Kriging(X, y)
    e = rng.uniform(-4, 0.5, (40, kk)); p = rng.uniform(1, 2, (40, kk)); p[:5] = 2.0; e[:5] = -4.5   # some ill-conditioned rows
    v = kr.likelihood_batch(e, p)
    g = kr.grid_search(3, 2)
    kr.update_param(10. ** e[7], p[7])
    yy, mm = kr.predict_batch(rng.uniform(0, 10, (37, kk)), return_mse=True)
    kr.update_param(10. ** e[0], p[0])                      # ill-conditioned fit -> double-double weights
    print(n, kk, np.isfinite(v).sum(), g[1], float(yy[0]), kr._fit['info'], kr.calc_likelihood())
r = RBF(rng.uniform(0, 10, (130, 3)), rng.uniform(0, 1, 130))
for name in ('gaus', 'mq', 'lin'):
    r.update_param(0.7, name)
    print(name, r.predict([1., 2., 3.]))
print(device.mesh_points([0., 0.], [1., 2.], [7, 5]).sum().item())
L = _cabi.lib()
xq = np.ascontiguousarray(rng.uniform(-2, 12, (5000, 2))); out = np.empty(5000); mse = np.empty(5000); sc = np.empty(4)
th, pp = np.array([0.05, 0.05]), np.array([1.9, 1.9])
Xc, yc = np.ascontiguousarray(X5), np.ascontiguousarray(y5)
print(L.sm_krig_fit_predict_host(Xc.ctypes.data, yc.ctypes.data, 36, 2, th.ctypes.data, pp.ctypes.data, xq.ctypes.data, 5000,
                                 out.ctypes.data, mse.ctypes.data, sc.ctypes.data), out[:2])
torch.cuda.synchronize()
print('done')
