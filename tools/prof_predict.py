"""ncu driver: prediction kernels (C5 shape) and RBF evaluation (C4 shape) at reduced query counts."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch  # noqa: E402
from mylibs.kriging import Kriging  # noqa: E402
from mylibs.rbf import RBF  # noqa: E402
from sm_b200 import device, workloads  # noqa: E402

X5, y5 = workloads.c5_kriging_r3()
k5 = Kriging(X5, y5)
k5.update_param([0.05, 0.05], [1.9, 1.9])
xq = device.to_dev(workloads.c5_query_mesh(2000))
for _ in range(2):
    k5.predict_device(xq, True)
    k5.predict_device(xq, False)
X4, y4 = workloads.c4_rbf4096()
r4 = RBF(X4, y4)
r4.update_param(1.0, 'gaus')
xq4 = device.to_dev(workloads.c4_queries(500000))
for _ in range(2):
    r4.predict_device(xq4)
torch.cuda.synchronize()
print('ok')
