"""ncu driver: one list-mode likelihood launch at large n (whatever kernel the environment selects)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from sm_b200 import _cabi, device, workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
C = int(sys.argv[3]) if len(sys.argv) > 3 else 592
X, y = workloads.c3_lhs1024(n=n, k=k)
e, p = workloads.c3_candidates(C, k)
L = _cabi.lib()
dX, dy, th, pd = device.to_dev(X), device.to_dev(y), device.to_dev(10. ** e), device.to_dev(p)
out = device.empty(C)
f = lambda: _cabi.check(L.sm_krig_nll_batch(device.ptr(dX), device.ptr(dy), n, k, device.ptr(th), device.ptr(pd), C, 0, device.ptr(out), None, None, device.stream()))
f()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
f()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
fl = (n ** 3 / 3 + (k + 2) * n ** 2) * C
print(f'n={n} k={k} C={C}: {ms:.2f} ms, {C / ms * 1e3:.4g} evals/s, {fl / ms / 1e9:.2f} TFLOP/s')
