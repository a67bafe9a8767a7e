"""Per-phase warp-cycle accounting of the dataflow likelihood kernel (build with SMB_DEFINES=SMB_PHASE_TIMING).
Sums clock64 deltas over every warp of block 0."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from sm_b200 import _cabi, device, workloads  # noqa: E402

n, k, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
X, y = workloads.c3_lhs1024(n=n, k=k)
e, p = workloads.c3_candidates(C, k)
L = _cabi.lib()
dX, dy, th, pd = device.to_dev(X), device.to_dev(y), device.to_dev(10. ** e), device.to_dev(p)
out = device.empty(C)
f_list = lambda: _cabi.check(L.sm_krig_nll_batch(device.ptr(dX), device.ptr(dy), n, k, device.ptr(th), device.ptr(pd), C, 0, device.ptr(out), None, None, device.stream()))
if len(sys.argv) > 4 and sys.argv[4] == 'grid':  # C = T^k P^k candidates sharing the |dx|^p tables
    T, P = int(sys.argv[5]), int(sys.argv[6])
    C = (T * P) ** k
    ta, pa = device.to_dev(10. ** np.linspace(-2, 0.5, T)), device.to_dev(np.linspace(1., 2., P))
    out = device.empty(C)
    f = lambda: _cabi.check(L.sm_krig_nll_grid(device.ptr(dX), device.ptr(dy), n, k, device.ptr(ta), T, device.ptr(pa), P, 0, C,
                                               device.ptr(out), None, None, device.stream()))
else:
    f = f_list
f(); torch.cuda.synchronize()
buf = np.zeros(16, dtype=np.uint64)
L.smb_flow_phase_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.smb_flow_phase_read(buf.ctypes.data, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); f(); b.record(); torch.cuda.synchronize()
L.smb_flow_phase_read(buf.ctypes.data, 1)
names = ['grab/decode', 'P wait rowdone', 'P k-loop', 'P wait E', 'P load+cholesky', 'P wait W slot', 'P inverse', 'U wait (rows | P)', 'U k-loop',
         'U wait E', 'U rmw+solve+store', 'end barrier', 'E compute', 'E signal']
tot = float(buf[:14].sum())
print(f'n={n} k={k} C={C} cfg={os.environ.get("SMB_FLOW_CFG", "default")}: {a.elapsed_time(b):.2f} ms; block 0 warp-cycles by phase (all warps summed): total {tot/1e6:.2f} M')
for i, nm in enumerate(names):
    print(f'  {nm:16s} {buf[i]/1e3:12.1f} k  {100*buf[i]/tot:5.1f} %')
