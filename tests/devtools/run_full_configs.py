"""BASELINE.json configs 3, 4, 5 at their FULL sizes on one GPU, with spot checks against the oracle.
Writes gpurun_out/full_configs.json (copied to profiles/ by hand)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from mylibs.kriging import Kriging  # noqa: E402
from mylibs.rbf import RBF  # noqa: E402
from oracle import numpy_oracle as orc  # noqa: E402
from sm_b200 import device, workloads  # noqa: E402

out = {}


def ev_ms(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b), r


# ---- C3: n = 1024, k = 8, 1e5 explicit candidates
X, y = workloads.c3_lhs1024()
e, p = workloads.c3_candidates(100000, 8)
k3 = Kriging(X, y)
k3.likelihood_batch(e[:592], p[:592])
t0 = time.perf_counter()
nll, bv, bi = k3.likelihood_batch(e, p, return_best=True)
wall = time.perf_counter() - t0
assert bi == orc.first_min_vec(nll) and nll[bi] == bv
sel = [bi, 0, 50000, 99999]
ref = np.array([orc.nll_chol(X, y, 10. ** e[c], p[c]) for c in sel])
err = float(np.max(np.abs(nll[sel] - ref) / np.maximum(1, np.abs(ref))))
out['c3'] = dict(n=1024, k=8, candidates=100000, wall_s_incl_copies=wall, evals_per_s=100000 / wall,
                 tflops_fma_only=(1024 ** 3 / 3 + 10 * 1024 ** 2) * 1e5 / wall / 1e12, argmin=int(bi), best_nll=float(bv),
                 nonfinite=int(np.sum(~np.isfinite(nll))), spot_check_rel_err_vs_oracle=err,
                 cond_at_argmin=orc.cond2(orc.corr_matrix(X, 10. ** e[bi], p[bi])))
print('c3', out['c3'], flush=True)

# ---- C4: gaussian RBF n = 4096, 1e7 evaluation points
X4, y4 = workloads.c4_rbf4096()
r4 = RBF(X4, y4)
r4.update_param(1.0, 'gaus')
torch.cuda.synchronize()
t0 = time.perf_counter()
r4.update_param(1.0, 'gaus')
torch.cuda.synchronize()
fit_s = time.perf_counter() - t0
xq = workloads.c4_queries(10 ** 7)
xq_d = device.to_dev(xq)
r4.predict_device(xq_d[:1000])
ms, yq = ev_ms(lambda: r4.predict_device(xq_d))
yq = yq.cpu().numpy()
ref = orc.rbf_predict(X4, r4.get_coeff(), 2, 1.0, xq[::400000])
out['c4'] = dict(n=4096, k=3, queries=10 ** 7, weight_solve_s=fit_s, eval_ms=ms, evals_per_s=1e7 / ms * 1e3,
                 spot_check_rel_err_vs_oracle=float(np.max(np.abs(yq[::400000] - ref) / np.maximum(1, np.abs(ref)))),
                 interpolation_max_abs=float(np.max(np.abs(r4.predict_batch(X4[::64]) - y4[::64]))))
print('c4', out['c4'], flush=True)
del xq_d, yq

# ---- C5: kriging_r3 model, 1e8 mesh queries, value + MSE
X5, y5 = workloads.c5_kriging_r3()
k5 = Kriging(X5, y5)
k5.update_param([0.05, 0.05], [1.9, 1.9])
m = 10000
t0 = time.perf_counter()
xq5 = workloads.c5_query_mesh(m)
gen_s = time.perf_counter() - t0
t0 = time.perf_counter()
xq5_d = device.to_dev(xq5)
torch.cuda.synchronize()
h2d_s = time.perf_counter() - t0
k5.predict_device(xq5_d[:1000], True)
ms, (yy, mse) = ev_ms(lambda: k5.predict_device(xq5_d, True))
ms_mean, _ = ev_ms(lambda: k5.predict_device(xq5_d, False))
idx = np.arange(0, m * m, 99991)
yr, mr = orc.predict_chol(X5, y5, [0.05, 0.05], [1.9, 1.9], xq5[idx], True)
yy_s, mse_s = yy[torch.from_numpy(idx).cuda()].cpu().numpy(), mse[torch.from_numpy(idx).cuda()].cpu().numpy()
out['c5'] = dict(n=36, k=2, queries=m * m, predict_value_mse_ms=ms, predictions_per_s=m * m / ms * 1e3,
                 predict_mean_only_ms=ms_mean, h2d_s=h2d_s, hbm_gbs=m * m * 32 / ms / 1e6,
                 spot_check_value_rel_err=float(np.max(np.abs(yy_s - yr) / np.maximum(1, np.abs(yr)))),
                 spot_check_mse_abs_err_over_sigma2=float(np.max(np.abs(mse_s - mr)) / k5._fit['sigma2']),
                 mse_min=float(mse.min().item()), checksum_value=float(yy.sum().item()))
print('c5', out['c5'], flush=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'full_configs.json'), 'w'), indent=1)
