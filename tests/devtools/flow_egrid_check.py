"""Dev tool: grid-mode likelihood of the dataflow kernel with product-form tables (default) against the exp-per-entry form
(SMB_FLOW_EGRID_OFF=1) and the oracle; throughput of both.  usage: flow_egrid_check.py [n,k,T,P ...]"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))

if len(sys.argv) > 1 and sys.argv[1] == 'worker':
    import torch
    from mylibs.kriging import Kriging
    from sm_b200 import workloads
    out = {}
    for spec in sys.argv[2:]:
        n, k, T, P = (int(v) for v in spec.split(','))
        X, y = workloads.c3_lhs1024(n=n, k=k)
        kr = Kriging(X * 10, y)
        bv, bi, nll = kr.likelihood_grid(T, P)
        g = nll.cpu().numpy()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = 1e30
        for _ in range(3):
            torch.cuda.synchronize()
            a.record()
            kr.likelihood_grid(T, P, return_nll=False)
            b.record()
            torch.cuda.synchronize()
            ms = min(ms, a.elapsed_time(b))
        out[spec] = dict(nll=[None if not np.isfinite(v) else float(v) for v in g], ms=ms, bi=int(bi))
    print(json.dumps(out))
    sys.exit(0)

from oracle import numpy_oracle as orc  # noqa: E402
from sm_b200 import workloads  # noqa: E402
specs = sys.argv[1:] or ['200,2,20,10', '256,2,20,10', '256,3,6,4', '300,4,4,3', '512,2,12,10']
res = {}
for flag in ('0', '1'):
    r = subprocess.run([sys.executable, os.path.abspath(__file__), 'worker', *specs], env=dict(os.environ, SMB_FLOW_EGRID_OFF=flag),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res[flag] = json.loads(r.stdout.strip().split('\n')[-1])
rng = np.random.default_rng(1)
for spec in specs:
    n, k, T, P = (int(v) for v in spec.split(','))
    a = np.array([np.nan if v is None else v for v in res['0'][spec]['nll']])
    b = np.array([np.nan if v is None else v for v in res['1'][spec]['nll']])
    m = np.isfinite(a) & np.isfinite(b)
    d = np.abs(a[m] - b[m]) / np.maximum(1, np.abs(b[m]))
    X, y = workloads.c3_lhs1024(n=n, k=k)
    eg, pg = orc.grid_candidates(k, T, P)
    idx = rng.choice(len(a), size=6, replace=False)
    ref = np.array([orc.nll_chol(X * 10, y, 10. ** eg[c], pg[c]) for c in idx])
    cond = np.array([orc.cond2(orc.corr_matrix(X * 10, 10. ** eg[c], pg[c])) for c in idx])
    mo = np.isfinite(ref) & np.isfinite(a[idx]) & (cond < 1e9)
    eo = np.max(np.abs(a[idx][mo] - ref[mo]) / np.maximum(1, np.abs(ref[mo])) / np.maximum(1.0, cond[mo] / 5e6)) if mo.any() else 0.0
    C = len(a)
    fl = n ** 3 / 3 + (k + 2) * n ** 2
    print(f'n={n} k={k} grid {T}^{k} x {P}^{k} = {C}: product form {res["0"][spec]["ms"]:.2f} ms ({fl * C / res["0"][spec]["ms"] / 1e9:.2f} TFLOP/s) | '
          f'exp form {res["1"][spec]["ms"]:.2f} ms ({fl * C / res["1"][spec]["ms"] / 1e9:.2f}) | finite {int(np.isfinite(a).sum())}/{int(np.isfinite(b).sum())} '
          f'median rel diff {np.median(d):.1e} max {d.max():.1e} | vs oracle ({int(mo.sum())} well-conditioned samples) err/tol-scale {eo:.1e} | '
          f'argmin {res["0"][spec]["bi"]} / {res["1"][spec]["bi"]}', flush=True)
    assert eo < 1e-9
