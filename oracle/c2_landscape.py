"""Map the ill-conditioned part of the two full C2 grids on the CPU (test infrastructure; uses oracle/quad_oracle.c).

For each response: fp64 Cholesky nll + smallest pivot for all 1e6 candidates, then the binary128 value for every
candidate whose smallest pivot is below --tau (or whose Cholesky failed).  Writes /tmp/land/<resp>.npz.
"""
import ctypes, os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
from oracle import numpy_oracle as orc
from sm_b200 import workloads as wl

lib = ctypes.CDLL(os.path.join(ROOT, 'oracle', '_build', 'libquad_oracle.so'))
dp = ctypes.POINTER(ctypes.c_double)
def P(a): return a.ctypes.data_as(dp)

def run(resp, tau):
    X = wl.halton_plan(64, 2, [(0., 10.), (0., 10.)])
    f = wl.f_3d if resp == 'f3d' else wl.fb
    y = np.array([f(x) for x in X])
    e, q = orc.grid_candidates(2, 40, 25)
    th = np.ascontiguousarray(10. ** e); q = np.ascontiguousarray(q)
    C = th.shape[0]
    nll = np.empty(C); mp = np.empty(C)
    t0 = time.time()
    lib.f64_nll_batch(P(X), P(y), 64, 2, P(th), P(q), ctypes.c_long(C), P(nll), P(mp))
    print(resp, 'f64 pass', round(time.time() - t0, 1), 's', flush=True)
    sel = np.flatnonzero(~(mp > tau))
    print('candidates with min pivot <=', tau, ':', len(sel), ' chol fails:', int(np.isinf(nll).sum()), flush=True)
    qn = np.empty(len(sel)); qmp = np.empty(len(sel)); qr = np.empty(len(sel)); qrmp = np.empty(len(sel))
    ths = np.ascontiguousarray(th[sel]); qs = np.ascontiguousarray(q[sel])
    t0 = time.time()
    lib.quad_nll_batch(P(X), P(y), 64, 2, P(ths), P(qs), ctypes.c_long(len(sel)), P(qn), P(qmp), None, None)
    lib.quad_nll_rounded_batch(P(X), P(y), 64, 2, P(ths), P(qs), ctypes.c_long(len(sel)), P(qr), P(qrmp))
    print('quad pass', round(time.time() - t0, 1), 's', flush=True)
    np.savez_compressed(f'/tmp/land/{resp}.npz', nll=nll, mp=mp, sel=sel, qn=qn, qmp=qmp, qr=qr, qrmp=qrmp)

if __name__ == '__main__':
    tau = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-9
    run(sys.argv[1], tau)
