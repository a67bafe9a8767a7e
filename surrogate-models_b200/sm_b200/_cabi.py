"""ctypes binding of include/surrogate_b200.h.  Fails loudly when the CUDA library is missing -- there is
no CPU fallback anywhere in the product path."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'lib', 'libsurrogate_b200.so')

c_dp = ctypes.POINTER(ctypes.c_double)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_ip = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_d = ctypes.c_double

# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
SIGNATURES = {
    'sm_b200_version': (_i, []),
    'sm_b200_error_string': (ctypes.c_char_p, [_i]),
    'sm_b200_launch_count': (ctypes.c_longlong, []),
    'sm_b200_trim_pool': (_i, []),
    'sm_exp10': (_i, [_vp, _i64, _vp, _vp]),
    'sm_krig_npad': (_i, [_i]),
    'sm_krig_mop_doubles': (ctypes.c_size_t, [_i]),
    'sm_krig_pivot_tau': (_d, []),
    'sm_krig_rescue_stats': (_i, [_vp, _vp, _i]),
    'sm_krig_nll_grid': (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _i, _i64, _i64, _vp, _vp, _vp, _vp]),
    'sm_krig_nll_grid_host': (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _i, _i64, _i64, _vp, _vp, _vp]),
    'sm_krig_loo_grid': (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    'sm_krig_nll_batch': (_i, [_vp, _vp, _i, _i, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    'sm_krig_nll_batch_host': (_i, [_vp, _vp, _i, _i, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    'sm_krig_fit': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'sm_krig_predict': (_i, [_vp, _i, _i, _vp, _vp, _vp, _d, _d, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'sm_krig_fit_predict_host': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'sm_mesh_points': (_i, [_vp, _vp, _vp, _i, _i64, _i64, _vp, _vp]),
    'sm_krig_predict_mesh': (_i, [_vp, _i, _i, _vp, _vp, _vp, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    'sm_krig_fit_predict_mesh_host': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    'sm_rbf_fit': (_i, [_vp, _vp, _i, _i, _i, _d, _vp, _vp, _vp]),
    'sm_rbf_predict': (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _i64, _vp, _vp]),
    'sm_rbf_fit_host': (_i, [_vp, _vp, _i, _i, _i, _d, _vp, _vp]),
    'sm_rbf_predict_host': (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _i64, _vp]),
}

_lib = None


class SurrogateB200Error(RuntimeError):
    pass


def lib():
    """The loaded library (built by surrogate-models_b200/build.py).  Raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SurrogateB200Error(
                f'{LIB_PATH} not found: build it with `python surrogate-models_b200/build.py` '
                '(there is no CPU fallback)')
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().sm_b200_error_string(int(rc)).decode()
        raise SurrogateB200Error(f'surrogate_b200 call failed ({rc}): {msg}')


def hptr(a):
    """Host pointer of a C-contiguous float64 / int64 / int32 numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def rescue_stats(reset=False):
    """(candidates re-evaluated in double-double, of those finite) since the last reset on the current device."""
    a, b = ctypes.c_int64(0), ctypes.c_int64(0)
    check(lib().sm_krig_rescue_stats(ctypes.byref(a), ctypes.byref(b), 1 if reset else 0))
    return int(a.value), int(b.value)


def launch_count():
    return int(lib().sm_b200_launch_count())
