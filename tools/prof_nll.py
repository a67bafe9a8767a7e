"""Small driver for ncu: one C2-shaped grid likelihood launch (n=64, k=2) over a slice of candidates."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from mylibs.kriging import Kriging  # noqa: E402
from sm_b200 import workloads  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
begin = int(sys.argv[3]) if len(sys.argv) > 3 else 400000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
if n == 64:
    X, y = workloads.c2_halton64()
else:
    X, y = workloads.c3_lhs1024(n=n, k=2)
    X = X * 10
k = Kriging(X, y)
for _ in range(2):
    bv, bi, _ = k.likelihood_grid(40, 25, begin, count, return_nll=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
bv, bi, _ = k.likelihood_grid(40, 25, begin, count, return_nll=True)
b.record()
torch.cuda.synchronize()
print(f'n={n} count={count} best {bv} @ {bi}: {a.elapsed_time(b):.3f} ms -> {count / a.elapsed_time(b) * 1e3:.4g} evals/s')
