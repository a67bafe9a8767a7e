"""Checks of csrc/dd.cuh on the host (tools/dd_check.cpp): exp_cr / pow_cr against glibc and against mpmath, and the
double-double likelihood against oracle/quad_oracle.c on the ill-conditioned candidates of the full C2 grid."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
lib = ctypes.CDLL('/tmp/libdd_check.so')
ql = ctypes.CDLL(os.path.join(ROOT, 'oracle', '_build', 'libquad_oracle.so'))
dp = ctypes.POINTER(ctypes.c_double)
P = lambda a: a.ctypes.data_as(dp)

def check_funcs():
    import mpmath as mp
    mp.mp.prec = 200
    rng = np.random.default_rng(0)
    x = -rng.uniform(0, 40, 400000) * rng.uniform(0, 1, 400000) ** 2
    out = np.empty_like(x); lib.exp_cr_vec(P(x), len(x), P(out))
    g = np.exp(x)
    print('exp: dd vs glibc mismatches', int((out != g).sum()), 'of', len(x))
    bad = np.flatnonzero(out != g)[:2000]
    nd = ng = 0
    for i in bad:
        t = float(mp.exp(mp.mpf(float(x[i]))))
        nd += out[i] != t; ng += g[i] != t
    print('   among', len(bad), 'mismatches: dd wrong', nd, ' glibc wrong', ng)
    xs = rng.uniform(0, 10, 400000); ps = rng.uniform(1, 2, 400000)
    ps[::4] = np.linspace(1, 2, 25)[rng.integers(0, 25, len(ps[::4]))]
    out = np.empty_like(xs); lib.pow_cr_vec(P(xs), P(ps), len(xs), P(out))
    g = xs ** ps
    print('pow: dd vs glibc mismatches', int((out != g).sum()), 'of', len(xs))
    bad = np.flatnonzero(out != g)[:2000]
    nd = ng = 0
    for i in bad:
        t = float(mp.power(mp.mpf(float(xs[i])), mp.mpf(float(ps[i]))))
        nd += out[i] != t; ng += g[i] != t
    print('   among', len(bad), 'mismatches: dd wrong', nd, ' glibc wrong', ng)
    # spot check of dd against mpmath where they agree with glibc is implied; random direct check too
    idx = rng.integers(0, len(x), 3000)
    print('exp direct: dd wrong in', sum(out_i != float(mp.exp(mp.mpf(float(xi)))) for out_i, xi in zip(np.array([lib_exp(x[i]) for i in idx]), x[idx])), 'of 3000')

def lib_exp(v):
    a = np.array([v]); o = np.empty(1); lib.exp_cr_vec(P(a), 1, P(o)); return o[0]

def check_c2():
    from oracle import numpy_oracle as orc
    from sm_b200 import workloads as wl
    X = wl.halton_plan(64, 2, [(0., 10.), (0., 10.)])
    e, q = orc.grid_candidates(2, 40, 25)
    for r in ('f3d', 'fb'):
        z = np.load(f'/tmp/land/{r}.npz'); sel = z['sel']
        f = wl.f_3d if r == 'f3d' else wl.fb
        y = np.array([f(x) for x in X])
        th = np.ascontiguousarray(10. ** e[sel]); qq = np.ascontiguousarray(q[sel])
        nll = np.empty(len(sel)); mpv = np.empty(len(sel))
        lib.dd_nll_batch(P(X), P(y), 64, 2, P(th), P(qq), ctypes.c_long(len(sel)), P(nll), P(mpv))
        qr, qrmp = z['qr'], z['qrmp']
        same_set = np.array_equal(np.isfinite(nll), np.isfinite(qr))
        fin = np.isfinite(qr) & np.isfinite(nll)
        print(r, 'acceptance identical to binary128 on glibc entries:', same_set, ' finite', int(fin.sum()),
              ' max |dd - quad|', np.abs(nll[fin] - qr[fin]).max(), ' argmin', int(sel[np.argmin(nll)]), nll.min())
        # entries: how many differ from glibc's
        nd = 0
        for c in sel[:40]:
            R = np.empty((64, 64)); lib.cr_entries(P(X), 64, 2, P(np.ascontiguousarray(10. ** e[c])), P(np.ascontiguousarray(q[c])), P(R))
            nd += int((R != orc.corr_matrix(X, 10. ** e[c], q[c])).sum())
        print('   entries differing from numpy/glibc over 40 matrices:', nd, 'of', 40 * 4096)

if __name__ == '__main__':
    check_funcs()
    check_c2()
