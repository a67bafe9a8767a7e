"""RBF weight-solve wall time, n = 4096 Halton samples (BASELINE config 4): gaussian (cooperative Cholesky) and mq (blocked LU)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import numpy as np, torch
from mylibs.rbf import RBF
from sm_b200 import workloads, _cabi
for n in (1024, 4096):
    X, y = workloads.c4_rbf4096(n)
    for name, a in (('gaus', 1.0), ('mq', 0.5), ('lin', 1.0)):
        r = RBF(X, y); r.update_param(a, name); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            n0 = _cabi.launch_count(); t0 = time.perf_counter(); r.update_param(a, name); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(f'n={n} {name}: fit {min(ts)*1e3:.2f} ms, {_cabi.launch_count()-n0} launches')
