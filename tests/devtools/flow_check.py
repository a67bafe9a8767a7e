"""Dev tool: parity + timing of the n > 128 likelihood kernels (dataflow kernel vs. the macro-panel kernel).

  python tests/devtools/flow_check.py                 # driver: runs every configuration in a subprocess (the kernel choice
                                             # is read from the environment once per process) and prints a table
  python tests/devtools/flow_check.py worker <json>   # one configuration
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))


def worker(spec):
    import torch
    from sm_b200 import _cabi, device, workloads
    from oracle import numpy_oracle as orc
    L = _cabi.lib()
    out = {'parity': [], 'timing': []}
    rng = np.random.default_rng(5)

    def batch(X, y, th, p):
        n, k = X.shape
        C = th.shape[0]
        dX, dy, dth, dp = device.to_dev(X), device.to_dev(y), device.to_dev(th), device.to_dev(p)
        o = device.empty(C)
        f = lambda: _cabi.check(L.sm_krig_nll_batch(device.ptr(dX), device.ptr(dy), n, k, device.ptr(dth), device.ptr(dp),
                                                    C, 0, device.ptr(o), None, None, device.stream()))
        return f, o

    for (n, k, C) in spec.get('parity', []):
        X, y = workloads.c3_lhs1024(n=n, k=k)
        e, p = workloads.c3_candidates(C, k)
        th = 10. ** e
        if C > 3:
            p[1] = 2.0
            p[2] = 1.0
            th[3] = 1e-9  # not SPD numerically -> +inf
        f, o = batch(X, y, th, p)
        f()
        torch.cuda.synchronize()
        got = o.cpu().numpy()
        ref = np.array([orc.nll_chol(X, y, th[c], p[c]) for c in range(C)])
        fin = np.isfinite(ref) & np.isfinite(got)
        err = float(np.max(np.abs(got[fin] - ref[fin]) / np.maximum(1, np.abs(ref[fin])))) if fin.any() else -1.0
        out['parity'].append({'n': n, 'k': k, 'C': C, 'relerr': err, 'nonfinite_match': bool(np.array_equal(np.isfinite(ref), np.isfinite(got))),
                              'nfinite': int(fin.sum()), 'got0': float(got[0]), 'ref0': float(ref[0]),
                              'each': [float(abs(g_ - r_) / max(1, abs(r_))) if np.isfinite(g_) and np.isfinite(r_) else float(np.isfinite(g_)) - float(np.isfinite(r_)) for g_, r_ in zip(got, ref)]})
    for (n, k, T, P) in spec.get('grid', []):
        X = rng.uniform(0, 10, (n, k))
        y = np.sin(X.sum(axis=1)) + 0.1 * X[:, 0] ** 2
        th_axis = np.linspace(-2, 0.5, T)
        p_axis = np.linspace(1., 2., P)
        Cn = (T ** k) * (P ** k)
        dX, dy = device.to_dev(X), device.to_dev(y)
        dta, dpa = device.to_dev(10. ** th_axis), device.to_dev(p_axis)
        o = device.empty(Cn)
        _cabi.check(L.sm_krig_nll_grid(device.ptr(dX), device.ptr(dy), n, k, device.ptr(dta), T, device.ptr(dpa), P, 0, Cn,
                                       device.ptr(o), None, None, device.stream()))
        torch.cuda.synchronize()
        got = o.cpu().numpy()
        ref = np.empty(Cn)
        for c in range(Cn):
            it, ip = c % (T ** k), c // (T ** k)
            th = np.array([10. ** th_axis[(it // T ** d) % T] for d in range(k)])
            pp = np.array([p_axis[(ip // P ** d) % P] for d in range(k)])
            ref[c] = orc.nll_chol(X, y, th, pp)
        fin = np.isfinite(ref) & np.isfinite(got)
        err = float(np.max(np.abs(got[fin] - ref[fin]) / np.maximum(1, np.abs(ref[fin])))) if fin.any() else -1.0
        out['parity'].append({'n': n, 'k': k, 'grid': [T, P], 'relerr': err, 'nfinite': int(fin.sum()),
                              'nonfinite_match': bool(np.array_equal(np.isfinite(ref), np.isfinite(got)))})
    for (n, k, C) in spec.get('timing', []):
        X, y = workloads.c3_lhs1024(n=n, k=k)
        e, p = workloads.c3_candidates(C, k)
        f, o = batch(X, y, 10. ** e, p)
        f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            a.record()
            f()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        fl = (n ** 3 / 3 + (k + 2) * n ** 2) * C
        out['timing'].append({'n': n, 'k': k, 'C': C, 'ms': best, 'evals_per_s': C / best * 1e3, 'tflops': fl / best / 1e9,
                              'checksum': float(np.nansum(o.cpu().numpy()))})
    print('RESULT ' + json.dumps(out), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == 'worker':
        worker(json.loads(sys.argv[2]))
        return
    quick = '--quick' in sys.argv
    parity = [(136, 2, 8), (200, 3, 8), (256, 8, 8), (300, 8, 8), (520, 4, 8), (1024, 8, 6)]
    grid = [(200, 2, 3, 2)]
    timing = [(256, 8, 9472), (512, 8, 2368), (1024, 8, 1184), (2048, 8, 296)]
    if quick:
        timing = [(256, 8, 2368), (512, 8, 1184), (1024, 8, 592)]
    configs = []

    for cfg in ('42', '63'):
        configs.append((f'flow cfg {cfg}', {'SMB_FLOW_CFG': cfg, 'SMB_FLOW_MIN_NT': '16'}, False))

    for name, env, with_parity in configs:
        spec = {'timing': timing}
        if with_parity:
            spec['parity'] = parity
            spec['grid'] = grid
        e = dict(os.environ)
        e.update(env)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), 'worker', json.dumps(spec)], env=e, capture_output=True,
                               text=True, timeout=300)
        except subprocess.TimeoutExpired:
            print(f'=== {name}: TIMEOUT', flush=True)
            continue
        print(f'=== {name} (exit {r.returncode}, {time.time() - t0:.0f} s)', flush=True)
        res = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
        if not res:
            print(r.stdout[-2000:], r.stderr[-3000:], flush=True)
            continue
        out = json.loads(res[0][7:])
        for p_ in out['parity']:
            print('  parity', json.dumps(p_), flush=True)
        for t_ in out['timing']:
            print(f"  timing n={t_['n']:5d} k={t_['k']} C={t_['C']:6d}: {t_['ms']:9.2f} ms  {t_['evals_per_s']:.4g} evals/s  {t_['tflops']:.2f} TFLOP/s  checksum {t_['checksum']:.12g}", flush=True)


if __name__ == '__main__':
    main()
