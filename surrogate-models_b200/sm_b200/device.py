"""Device-memory plumbing: torch owns HBM buffers and streams, the C-ABI gets raw pointers.

Nothing here computes; a missing GPU or library raises (no CPU fallback)."""
import numpy as np
import torch

from . import _cabi


def require_cuda():
    if not torch.cuda.is_available():
        raise _cabi.SurrogateB200Error('surrogate-models_b200 needs a CUDA device (B200, sm_100a); '
                                       'there is no CPU fallback')
    _cabi.lib()
    return torch.device('cuda', torch.cuda.current_device())


def to_dev(a, dtype=torch.float64):
    """numpy / sequence -> contiguous device tensor (pageable host memory, synchronous copy)."""
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    npdt = np.float64 if dtype == torch.float64 else np.int64
    return torch.from_numpy(np.ascontiguousarray(a, dtype=npdt)).to(dev)


def empty(shape, dtype=torch.float64):
    return torch.empty(shape, dtype=dtype, device=require_cuda())


def ptr(t):
    return None if t is None else t.data_ptr()


def mesh_points(lo, hi, m, begin=0, count=None):
    """rows [begin, begin + count) of the full-factorial query mesh (last dimension fastest), generated on the device
    (sm_mesh_points); bit-identical to stacking numpy.linspace axes with meshgrid(indexing='ij')."""
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    m = np.ascontiguousarray(m, dtype=np.int64)
    k = lo.shape[0]
    total = int(np.prod(m))
    count = total - begin if count is None else int(count)
    out = empty((count, k))
    _cabi.check(_cabi.lib().sm_mesh_points(lo.ctypes.data, hi.ctypes.data, m.ctypes.data, k, int(begin), count, ptr(out),
                                           stream()))
    return out


def stream():
    return torch.cuda.current_stream().cuda_stream
