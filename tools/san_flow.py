"""compute-sanitizer driver: a few small launches of every likelihood kernel family."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
from mylibs.kriging import Kriging  # noqa: E402
from sm_b200 import workloads  # noqa: E402

for n, k, C in [(150, 3, 6), (163, 2, 6), (40, 2, 40), (100, 6, 8), (140, 2, 6)]:
    X, y = workloads.c3_lhs1024(n=n, k=k)
    e, p = workloads.c3_candidates(C, k)
    kr = Kriging(X * 10, y)
    v = kr.likelihood_batch(e, p)
    print(n, k, v[:3], flush=True)
    if k <= 2:
        print(kr.likelihood_grid(3, 2)[:2], flush=True)
