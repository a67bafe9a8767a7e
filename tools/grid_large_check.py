"""Dev tool: grid-mode likelihood throughput at large n (tile-packed |dx|^p tables, dataflow kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from mylibs.kriging import Kriging  # noqa: E402
from sm_b200 import workloads  # noqa: E402

for n, k, T, P in [(256, 2, 40, 5), (512, 2, 30, 4), (1024, 2, 20, 3), (1024, 3, 8, 3)]:
    X, y = workloads.c3_lhs1024(n=n, k=k)
    kr = Kriging(X * 10, y)
    C = (T ** k) * (P ** k)
    kr.likelihood_grid(T, P, return_nll=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    bv, bi, _ = kr.likelihood_grid(T, P, return_nll=False)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    fl = n ** 3 / 3 + (k + 2) * n ** 2
    _, _, nll = kr.likelihood_grid(T, P, return_nll=True)
    import numpy as np
    fin = float(np.isfinite(nll.cpu().numpy()).mean())
    print(f'finite fraction {fin:.3f};', end=' ')
    print(f'n={n} k={k} grid {T}x{P}: {C} candidates {ms:.1f} ms {C / ms * 1e3:.4g} evals/s {fl * C / ms / 1e9:.2f} TFLOP/s best {bv:.6f} @ {bi}', flush=True)
