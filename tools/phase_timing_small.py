"""Per-phase cycle accounting of the n <= 64 likelihood kernel (build with SMB_DEFINES=SMB_PHASE_TIMING)."""
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

X, y = workloads.c2_halton64()
k = Kriging(X, y)
L = _cabi.lib()
L.smb_small_phase_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
C = 1000000
k.likelihood_grid(40, 25, 0, C, return_nll=False)
torch.cuda.synchronize()
buf = np.zeros(8, dtype=np.uint64)
L.smb_small_phase_read(buf.ctypes.data, 1)
k.likelihood_grid(40, 25, 0, C, return_nll=False)
torch.cuda.synchronize()
L.smb_small_phase_read(buf.ctypes.data, 1)
ncand = C / (296 * 7)
names = ['decode', 'U: DMMA update loop', 'entries (tables, products)', 'park + rhs', 'D: diagonal factor', 'T: panel solve', 'epilogue']
tot = float(buf[:7].sum())
print(f'warp 0 of block 0: {ncand:.0f} candidates, {tot / ncand / 1e3:.1f}k cycles per candidate')
for nm, v in zip(names, buf[:7]):
    print(f'  {nm:28s} {float(v) / ncand / 1e3:8.2f}k cycles  {100 * float(v) / tot:5.1f} %')
