"""Dev tool (library built with SMB_DEFINES=SMB_PHASE_TIMING): cycles per phase of the CTA-per-candidate likelihood kernel,
CTA 0 only, owner of the column / the other warps.  usage: phase_timing_cta.py [n] [candidates]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from mylibs.kriging import Kriging  # noqa: E402
from sm_b200 import _cabi, workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
C = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
X, y = workloads.c3_lhs1024(n=n, k=2)
kr = Kriging(X * 10, y)
L = _cabi.lib()
L.smb_cta_phase_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
kr.likelihood_grid(40, 25, 0, C, return_nll=False)
torch.cuda.synchronize()
buf = np.zeros(16, dtype=np.uint64)
L.smb_cta_phase_read(buf.ctypes.data, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
kr.likelihood_grid(40, 25, 0, C, return_nll=False)
b.record()
torch.cuda.synchronize()
L.smb_cta_phase_read(buf.ctypes.data, 1)
NT = (n + 7) // 8
grid = min(C, 148 * 5)
cands = (C + grid - 1) // grid  # candidates of CTA 0
names = ["k-steps of column J", "C = R - acc, rhs", "diag (owner)", "entries + look-ahead", "wait: barrier 1 (W_J)", "solve, stores, K = J from regs", "-", "-"]

print(f'n={n} NT={NT} {C} candidates {a.elapsed_time(b):.3f} ms; CTA 0: {cands} candidates; cycles per block column')
for i, nm in enumerate(names):
    print(f'  {nm:28s} owner {buf[i] / (cands * NT):9.0f}   others (per warp) {buf[8 + i] / (cands * NT * 3):9.0f}')
print(f'  total                        owner {buf[:8].sum() / (cands * NT):9.0f}   others            {buf[8:].sum() / (cands * NT * 3):9.0f}')
