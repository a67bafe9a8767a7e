import os, sys, time
import numpy as np
ROOT = '/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch
from mylibs.kriging import Kriging
from sm_b200 import workloads
for n, k in [(256, 2), (512, 8), (1024, 8), (2048, 3)]:
    X, y = workloads.c3_lhs1024(n=n, k=k)
    e, p = workloads.c3_candidates(4, k)
    kr = Kriging(X, y)
    kr.update_param(10. ** e[0], p[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(1, 4):
        kr.update_param(10. ** e[i], p[i])
    dt = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    kr._run_fit(want_mse=True)
    dt_full = time.perf_counter() - t0
    print(f'n={n} k={k}: update_param {dt * 1e3:.2f} ms (weights only) vs {dt_full * 1e3:.2f} ms (full fit with L^-1)', flush=True)
