"""Per-phase cycle accounting of the large-n likelihood kernel (needs a build with SMB_DEFINES=SMB_PHASE_TIMING)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from mylibs.kriging import Kriging  # noqa: E402
from sm_b200 import _cabi, device, workloads  # noqa: E402

n, k, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
X, y = workloads.c3_lhs1024(n=n, k=k)
e, p = workloads.c3_candidates(C, k)
kr = Kriging(X, y)
L = _cabi.lib()
d = kr._dev()
th, pd = device.to_dev(10. ** e), device.to_dev(p)
out = device.empty(C)
f = lambda: _cabi.check(L.sm_krig_nll_batch(device.ptr(d['X']), device.ptr(d['y']), n, k, device.ptr(th), device.ptr(pd), C, 0, device.ptr(out), None, None, device.stream()))
f(); torch.cuda.synchronize()
buf = np.zeros((16, 8), dtype=np.uint64)
L.smb_phase_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.smb_phase_read(buf.ctypes.data, 1)
f(); torch.cuda.synchronize()
L.smb_phase_read(buf.ctypes.data, 1)
names = ['entries(0)', 'sync', 'bulk+rmw', 'sync after bulk', 'P1 (w0) / next entries (w1-7)', 'sync after P1', 'P3 panel solve', 'sync after P3']
cand = (C + 295) // 296 if C > 296 else 1
tot = buf[:8].sum(axis=0).astype(float)
print(f'n={n} k={k} C={C}: block 0 processed ~{cand} candidates; cycles per candidate by phase (warp 0 | warp 3 | warp 7), % of warp total')
for i, nm in enumerate(names):
    r = buf[i].astype(float)
    print(f'  {nm:34s} ' + ' | '.join(f'{r[w] / cand / 1e3:9.1f}k {100 * r[w] / tot[w]:5.1f}%' for w in (0, 3, 7)))
print('  total per candidate (k cycles):', ' | '.join(f'{tot[w] / cand / 1e3:9.1f}k' for w in (0, 3, 7)))
