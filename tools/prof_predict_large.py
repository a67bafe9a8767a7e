"""ncu / timing driver: value + MSE prediction at n > 128 (streamed-DMMA kernel; SMB_PREDICT_GENERIC=1 -> thread-per-query kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import numpy as np, torch
from mylibs.kriging import Kriging
from sm_b200 import device, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
X, y = workloads.c3_lhs1024(n=n, k=k)
kr = Kriging(X, y)
kr.update_param(np.full(k, 3.0), np.full(k, 1.6))
xq = device.to_dev(np.random.default_rng(0).uniform(0, 1, (nq, k)))
kr.predict_device(xq, True); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); kr.predict_device(xq, True); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
fl = (2 * n * k + 2 * n + n * n + 3 * n) * nq
print(f'n={n} k={k} nq={nq}: {ms:.2f} ms, {nq / ms * 1e3:.4g} predictions/s, {fl / ms / 1e9:.2f} TFLOP/s FMA-only ({fl / ms / 1e9 / 37.117:.1%} of peak)')
