"""Dev tool: likelihood throughput at 64 < n <= 144 (list and grid mode), whatever kernels the build selects."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from mylibs.kriging import Kriging  # noqa: E402
from sm_b200 import _cabi, device, workloads  # noqa: E402

L = _cabi.lib()
SIZES = [(128, 2, 20000), (136, 2, 20000), (144, 2, 20000), (144, 8, 20000), (152, 2, 20000)]
if len(sys.argv) > 1:  # n,k[;n,k...]
    SIZES = [(int(t.split(',')[0]), int(t.split(',')[1]), 200000 if int(t.split(",")[0]) <= 96 else 20000) for t in sys.argv[1].split(';')]
for n, k, C in SIZES:
    X, y = workloads.c3_lhs1024(n=n, k=k)
    e, p = workloads.c3_candidates(C, k)
    kr = Kriging(X * 10, y)
    d = kr._dev()
    th, pd = device.to_dev(10. ** e), device.to_dev(p)
    out = device.empty(C)
    f = lambda: _cabi.check(L.sm_krig_nll_batch(device.ptr(d['X']), device.ptr(d['y']), n, k, device.ptr(th), device.ptr(pd), C, 0,
                                                device.ptr(out), None, None, device.stream()))
    f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    fl = n ** 3 / 3 + (k + 2) * n ** 2
    if k > 2:
        print(f'n={n:4d} k={k}: list {C / ms * 1e3:.4g} evals/s {fl * C / ms / 1e9:.2f} TFLOP/s', flush=True)
        continue
    T, P = 20, 10
    bv, bi, _ = kr.likelihood_grid(T, P)
    torch.cuda.synchronize()
    a.record(); kr.likelihood_grid(T, P, return_nll=False); b.record(); torch.cuda.synchronize()
    Cg = (T ** k) * (P ** k)
    fl = n ** 3 / 3 + (k + 2) * n ** 2
    print(f'n={n:4d} k={k}: list {C / ms * 1e3:.4g} evals/s {fl * C / ms / 1e9:.2f} TFLOP/s | grid ({Cg} cand) {Cg / a.elapsed_time(b) * 1e3:.4g} evals/s '
          f'{fl * Cg / a.elapsed_time(b) / 1e9:.2f} TFLOP/s', flush=True)
