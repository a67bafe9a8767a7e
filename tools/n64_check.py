"""A/B of the statically unrolled npad = 64 kernel against the generic warp-per-candidate kernel: run twice (SMB_N64_OFF=0/1),
each run saves its nll vectors; the second run compares bit for bit and prints both timings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import numpy as np, torch
from mylibs.kriging import Kriging
from sm_b200 import workloads
tag = 'off' if os.environ.get('SMB_N64_OFF') == '1' else 'on'
X, y = workloads.c2_halton64()
k = Kriging(X, y)
out = {}
for _ in range(2):
    bv, bi, nll = k.likelihood_grid(40, 25)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); bv, bi, nll = k.likelihood_grid(40, 25); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
out['grid'] = nll.cpu().numpy()
print(f'[{tag}] grid 1e6: {min(ts):.3f} ms  best {bv} @ {bi}')
e, p = workloads.c3_candidates(200000, 2, seed=3)
ts = []
for _ in range(3):
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); v = k.likelihood_batch(e * 1.2 - 0.5, p); t1.record(); torch.cuda.synchronize(); ts.append(t0.elapsed_time(t1))
out['list'] = v
print(f'[{tag}] list 2e5 (incl. copies): {min(ts):.3f} ms')
X3 = np.random.default_rng(1).uniform(0, 10, (60, 3)); y3 = np.sin(X3.sum(axis=1))
k3 = Kriging(X3, y3)
_, _, n3 = k3.likelihood_grid(6, 3)      # MODE 2, k = 3, n = 60
out['k3'] = n3.cpu().numpy()
np.savez(f'/tmp/n64_{tag}.npz', **out)
other = f'/tmp/n64_{"on" if tag == "off" else "off"}.npz'
if os.path.exists(other):
    o = np.load(other)
    for key in out:
        print(key, 'bit-identical:', np.array_equal(out[key], o[key], equal_nan=True), 'finite', int(np.isfinite(out[key]).sum()))
