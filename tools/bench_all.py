"""Times every BASELINE config shape on one GPU (dev tool; bench.py is the contract benchmark)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from mylibs.kriging import Kriging  # noqa: E402
from mylibs.rbf import RBF  # noqa: E402
from sm_b200 import device, workloads  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def flops_like(n, k):
    return n ** 3 / 3 + (k + 2) * n ** 2


which = sys.argv[1:] or ['mid', 'c3', 'c4', 'c5']
if 'mid' in which:
    for n, k, C in ((128, 2, 20000), (256, 2, 4000), (256, 8, 4000), (512, 8, 1000)):
        X, y = workloads.c3_lhs1024(n=n, k=k)
        e, p = workloads.c3_candidates(C, k)
        kr = Kriging(X, y)
        th, pd = device.to_dev(10. ** e), device.to_dev(p)
        out = device.empty(C)
        from sm_b200 import _cabi
        L = _cabi.lib()
        d = kr._dev()
        f = lambda: _cabi.check(L.sm_krig_nll_batch(device.ptr(d['X']), device.ptr(d['y']), n, k, device.ptr(th), device.ptr(pd), C, 0, device.ptr(out), None, None, device.stream()))
        ms = timed(f)
        print(f'nll list n={n} k={k} C={C}: {ms:.2f} ms -> {C / ms * 1e3:.4g} evals/s, {flops_like(n, k) * C / ms / 1e9:.2f} TFLOP/s (FMA-only roofline 37.1)', flush=True)
if 'c3' in which:
    n, k, C = 1024, 8, 1200
    X, y = workloads.c3_lhs1024()
    e, p = workloads.c3_candidates(C, k)
    kr = Kriging(X, y)
    from sm_b200 import _cabi
    L = _cabi.lib()
    d = kr._dev()
    th, pd = device.to_dev(10. ** e), device.to_dev(p)
    out = device.empty(C)
    f = lambda: _cabi.check(L.sm_krig_nll_batch(device.ptr(d['X']), device.ptr(d['y']), n, k, device.ptr(th), device.ptr(pd), C, 0, device.ptr(out), None, None, device.stream()))
    ms = timed(f, reps=2)
    print(f'C3 nll list n=1024 k=8 C={C}: {ms:.1f} ms -> {C / ms * 1e3:.4g} evals/s, {flops_like(n, k) * C / ms / 1e9:.2f} TFLOP/s; 1e5 candidates would take {ms * 1e5 / C / 1e3:.1f} s', flush=True)
    t0 = time.time(); kr.update_param(10. ** e[0], p[0]); print(f'   fit n=1024: {time.time() - t0:.3f} s', flush=True)
    xq = device.to_dev(np.random.default_rng(0).uniform(0, 1, (200000, 8)))
    ms = timed(lambda: kr.predict_device(xq, False), reps=2)
    print(f'   predict mean n=1024 k=8 Nq=2e5: {ms:.1f} ms -> {2e5 / ms * 1e3:.4g} pred/s', flush=True)
if 'c4' in which:
    X, y = workloads.c4_rbf4096()
    r = RBF(X, y)
    t0 = time.time(); r.update_param(1.0, 'gaus'); torch.cuda.synchronize(); t1 = time.time() - t0
    t0 = time.time(); r.update_param(1.0, 'gaus'); torch.cuda.synchronize(); t2 = time.time() - t0
    print(f'C4 rbf fit n=4096 gaus: {t1:.3f} s first, {t2:.3f} s second', flush=True)
    Nq = 2000000
    xq = device.to_dev(workloads.c4_queries(Nq))
    ms = timed(lambda: r.predict_device(xq), reps=2)
    print(f'C4 rbf predict n=4096 Nq={Nq}: {ms:.1f} ms -> {Nq / ms * 1e3:.4g} pred/s ({Nq * 4096 / ms / 1e6:.1f} G kernel evals/s); 1e7 would take {ms * 5 / 1e3:.2f} s', flush=True)
    r2 = RBF(X[:1024], y[:1024])
    t0 = time.time(); r2.update_param(0.5, 'mq'); torch.cuda.synchronize(); print(f'   rbf fit n=1024 mq (LU): {time.time() - t0:.3f} s', flush=True)
if 'c5' in which:
    X, y = workloads.c5_kriging_r3()
    kr = Kriging(X, y)
    m = 3162
    xq = device.to_dev(workloads.c5_query_mesh(m))
    for th, p in (([0.05, 0.05], [1.9, 1.9]), ([0.05, 0.05], [2.0, 2.0])):
        kr.update_param(th, p)
        for mse in (False, True):
            ms = timed(lambda: kr.predict_device(xq, mse))
            print(f'C5 predict n=36 p={p[0]} mse={mse} Nq={m * m}: {ms:.2f} ms -> {m * m / ms * 1e3:.4g} pred/s; 1e8 would take {ms * 10:.0f} ms', flush=True)
    t0 = time.time(); kr.optimize(); print(f'   optimize(grid) n=36: {time.time() - t0:.3f} s total', flush=True)
