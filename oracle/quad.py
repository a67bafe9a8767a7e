"""ctypes front end of oracle/quad_oracle.c (binary128 / fp64 Cholesky likelihood) -- TEST INFRASTRUCTURE ONLY.

build() compiles the C file with gcc into oracle/_build/ (git-ignored; the .so travels to the GPU box with the
gpurun snapshot like the product's own library).  Used by tests/ and oracle/gen_golden.py, never by the product.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'quad_oracle.c')
LIB = os.path.join(HERE, '_build', 'libquad_oracle.so')
_dp = ctypes.POINTER(ctypes.c_double)
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(['gcc', '-O2', '-fopenmp', '-shared', '-fPIC', SRC, '-o', LIB, '-lquadmath', '-lm'], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _prep(X, y, theta, p):
    X = np.ascontiguousarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    y = np.ascontiguousarray(y, dtype=np.float64)
    theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
    return X, y, theta, p, X.shape[0], X.shape[1], theta.shape[0]


def nll_exact(X, y, theta, p):
    """binary128 likelihood from the fp64 inputs (entries, factorisation and logs in 113 bits): (nll, min_pivot)."""
    X, y, theta, p, n, k, C = _prep(X, y, theta, p)
    out, mp = np.empty(C), np.empty(C)
    lib().quad_nll_batch(_p(X), _p(y), n, k, _p(theta), _p(p), ctypes.c_long(C), _p(out), _p(mp), None, None)
    return out, mp


def nll_rounded(X, y, theta, p):
    """binary128 likelihood of the fp64-ROUNDED correlation matrix (entries through libm exactly as the reference
    forms them, kriging.py:70-77): (nll, min_pivot); +inf where that matrix is not positive definite."""
    X, y, theta, p, n, k, C = _prep(X, y, theta, p)
    out, mp = np.empty(C), np.empty(C)
    lib().quad_nll_rounded_batch(_p(X), _p(y), n, k, _p(theta), _p(p), ctypes.c_long(C), _p(out), _p(mp))
    return out, mp


def nll_f64(X, y, theta, p):
    """plain fp64 left-looking Cholesky likelihood: (nll, min_pivot)."""
    X, y, theta, p, n, k, C = _prep(X, y, theta, p)
    out, mp = np.empty(C), np.empty(C)
    lib().f64_nll_batch(_p(X), _p(y), n, k, _p(theta), _p(p), ctypes.c_long(C), _p(out), _p(mp))
    return out, mp
