"""Dev tool: randomised parity sweep of the likelihood kernels against the numpy oracle (sizes, dimensions, modes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
from mylibs.kriging import Kriging  # noqa: E402
from oracle import numpy_oracle as orc  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rng = np.random.default_rng(seed)
worst = 0.0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    n = int(rng.integers(9, 700))
    k = int(rng.integers(1, 9))
    X = rng.uniform(0, 10, (n, k))
    y = np.sin(X.sum(axis=1)) + 0.1 * X[:, 0] ** 2
    C = 5
    e = rng.uniform(-1.5, 0.3, (C, k))
    p = rng.uniform(1, 2, (C, k))
    p[0] = 2.0 if trial % 3 == 0 else p[0]
    kr = Kriging(X, y)
    got = kr.likelihood_batch(e, p)
    ref = np.array([orc.nll_chol(X, y, 10. ** e[c], p[c]) for c in range(C)])
    cond = np.array([orc.cond2(orc.corr_matrix(X, 10. ** e[c], p[c])) for c in range(C)])
    m = np.isfinite(ref) & np.isfinite(got) & (cond < 1e9)
    err = np.max(np.abs(got[m] - ref[m]) / np.maximum(1, np.abs(ref[m])) / np.maximum(1.0, cond[m] / 5e6)) if m.any() else 0.0
    bad_fin = np.isfinite(ref) & (cond < 1e9) & ~np.isfinite(got)
    worst = max(worst, err)
    line = f'list n={n:4d} k={k} ok={int(m.sum())}/{C} err/tol-scale {err:.2e}'
    if k <= 2 and n <= 400:
        T, P = 4, 3
        bv, bi, nll = kr.likelihood_grid(T, P)
        g = nll.cpu().numpy()
        eg, pg = orc.grid_candidates(k, T, P)
        idx = rng.choice(len(g), size=min(6, len(g)), replace=False)
        rg = np.array([orc.nll_chol(X, y, 10. ** eg[c], pg[c]) for c in idx])
        cg = np.array([orc.cond2(orc.corr_matrix(X, 10. ** eg[c], pg[c])) for c in idx])
        mg = np.isfinite(rg) & np.isfinite(g[idx]) & (cg < 1e9)
        eg_ = np.max(np.abs(g[idx][mg] - rg[mg]) / np.maximum(1, np.abs(rg[mg])) / np.maximum(1.0, cg[mg] / 5e6)) if mg.any() else 0.0
        worst = max(worst, eg_)
        line += f' | grid ok={int(mg.sum())}/{len(idx)} err {eg_:.2e}'
    print(line + (' NONFINITE-MISMATCH' if bad_fin.any() else ''), flush=True)
print('worst', worst)
assert worst < 1e-9
