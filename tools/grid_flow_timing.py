"""Dev tool: median-of-7 timing of sm_krig_nll_grid at dataflow-kernel sizes (whole call: tables + kernel + double-double pass).
usage: grid_flow_timing.py n,k,T,P [...]   (compare builds / SMB_FLOW_EGRID_OFF=1 / SMB_RESCUE_MAX_N=0 by environment)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from mylibs.kriging import Kriging  # noqa: E402
from sm_b200 import workloads  # noqa: E402

for spec in sys.argv[1:]:
    n, k, T, P = (int(v) for v in spec.split(','))
    X, y = workloads.c3_lhs1024(n=n, k=k)
    kr = Kriging(X * 10, y)
    kr.likelihood_grid(T, P, return_nll=False)
    ts = []
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(7):
        torch.cuda.synchronize()
        a.record()
        kr.likelihood_grid(T, P, return_nll=False)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    C = (T * P) ** k
    fl = n ** 3 / 3 + (k + 2) * n ** 2
    ms = float(np.median(ts))
    print(f'n={n} k={k} {C} candidates: median {ms:.2f} ms (min {min(ts):.2f}, max {max(ts):.2f})  {C / ms * 1e3:.4g} evals/s  {fl * C / ms / 1e9:.2f} TFLOP/s', flush=True)
