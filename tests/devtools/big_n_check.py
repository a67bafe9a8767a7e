import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch
from mylibs.kriging import Kriging
from oracle import numpy_oracle as orc
rng = np.random.default_rng(4096)
for n in (4096, 4104):
    X = rng.uniform(0, 1, (n, 3))
    y = np.sin(3 * X.sum(axis=1)) + X[:, 0] ** 2
    e = np.array([[1.0, 1.2, 0.9], [1.5, 1.0, 1.3]])
    p = np.array([[1.3, 1.6, 1.1], [1.0, 1.9, 1.5]])
    k = Kriging(X, y)
    t0 = time.time()
    nll = k.likelihood_batch(e, p)
    dt = time.time() - t0
    ref = np.array([orc.nll_chol(X, y, 10. ** e[c], p[c]) for c in range(2)])
    print(n, nll, ref, np.abs(nll - ref) / np.maximum(1, np.abs(ref)), f'{dt:.2f} s', flush=True)
