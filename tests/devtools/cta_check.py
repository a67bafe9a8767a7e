"""Dev tool: the CTA-per-candidate kernel (krig_nll_cta.cu) against the numpy oracle over its whole size range, both modes,
k = 1..4, plus throughput.  usage: cta_check.py [parity|speed|both] [n,n,...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from mylibs.kriging import Kriging  # noqa: E402
from oracle import numpy_oracle as orc  # noqa: E402
from sm_b200 import workloads  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'both'
sizes = [int(t) for t in sys.argv[2].split(',')] if len(sys.argv) > 2 else [81, 88, 96, 100, 104, 112, 120, 128, 129, 136, 144, 152, 160]
rng = np.random.default_rng(5)
worst = 0.0
if what in ('parity', 'both'):
    for n in sizes:
        for k in (1, 2, 3, 4):
            X = rng.uniform(0, 10, (n, k))
            y = np.sin(X.sum(axis=1)) + 0.1 * X[:, 0] ** 2
            C = 300  # more candidates than one wave of CTAs at the small sizes would need is not the point: reuse of a CTA is
            e = rng.uniform(-1.5, 0.3, (C, k))
            p = rng.uniform(1, 2, (C, k))
            p[::3] = 2.0
            kr = Kriging(X, y)
            got = kr.likelihood_batch(e, p)
            idx = rng.choice(C, size=8, replace=False)
            ref = np.array([orc.nll_chol(X, y, 10. ** e[c], p[c]) for c in idx])
            cond = np.array([orc.cond2(orc.corr_matrix(X, 10. ** e[c], p[c])) for c in idx])
            m = np.isfinite(ref) & np.isfinite(got[idx]) & (cond < 1e9)
            err = np.max(np.abs(got[idx][m] - ref[m]) / np.maximum(1, np.abs(ref[m])) / np.maximum(1.0, cond[m] / 5e6)) if m.any() else 0.0
            bad = np.isfinite(ref) & (cond < 1e9) & ~np.isfinite(got[idx])
            line = f'n={n:4d} k={k} list ok={int(m.sum())}/8 err {err:.2e}'
            worst = max(worst, err)
            T, P = (6, 3) if k <= 2 else (3, 2)
            bv, bi, nll = kr.likelihood_grid(T, P)
            g = nll.cpu().numpy()
            eg, pg = orc.grid_candidates(k, T, P)
            gi = rng.choice(len(g), size=min(6, len(g)), replace=False)
            rg = np.array([orc.nll_chol(X, y, 10. ** eg[c], pg[c]) for c in gi])
            cg = np.array([orc.cond2(orc.corr_matrix(X, 10. ** eg[c], pg[c])) for c in gi])
            mg = np.isfinite(rg) & np.isfinite(g[gi]) & (cg < 1e9)
            e2 = np.max(np.abs(g[gi][mg] - rg[mg]) / np.maximum(1, np.abs(rg[mg])) / np.maximum(1.0, cg[mg] / 5e6)) if mg.any() else 0.0
            worst = max(worst, e2)
            fin = g[np.isfinite(g)]
            first_min_ok = (not len(fin)) or (int(bi) == int(np.flatnonzero(g == fin.min())[0]))
            print(line + f' | grid ok={int(mg.sum())}/{len(gi)} err {e2:.2e} argmin_ok={first_min_ok}' + (' NONFINITE-MISMATCH' if bad.any() else ''),
                  flush=True)
            assert first_min_ok
    print('worst', worst)
    assert worst < 1e-9
if what in ('speed', 'both'):
    for n in sizes:
        for k in (2, 4):
            X, y = workloads.c3_lhs1024(n=n, k=k)
            C = 40000
            e, p = workloads.c3_candidates(C, k)
            kr = Kriging(X * 10, y)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            kr.likelihood_batch(e[:2000], p[:2000])
            th = torch.as_tensor(10. ** e, device='cuda')
            pp = torch.as_tensor(p, device='cuda')
            kr._likelihood_list_dev(th, pp)
            ms = 1e30
            for _ in range(5):
                torch.cuda.synchronize()
                a.record()
                kr._likelihood_list_dev(th, pp)
                b.record()
                torch.cuda.synchronize()
                ms = min(ms, a.elapsed_time(b))
            fl = n ** 3 / 3 + (k + 2) * n ** 2
            line = f'n={n:4d} k={k}: list {C / ms * 1e3:.4g} evals/s {fl * C / ms / 1e9:.2f} TFLOP/s'
            if k in (2, 4):
                T, P = (20, 10) if k == 2 else (5, 3)
                kr.likelihood_grid(T, P, return_nll=False)
                mg = 1e30
                for _ in range(5):
                    torch.cuda.synchronize()
                    a.record()
                    kr.likelihood_grid(T, P, return_nll=False)
                    b.record()
                    torch.cuda.synchronize()
                    mg = min(mg, a.elapsed_time(b))
                Cg = (T * P) ** k
                line += f' | grid ({Cg}) {Cg / mg * 1e3:.4g} evals/s {fl * Cg / mg / 1e9:.2f} TFLOP/s'
            print(line, flush=True)
