import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import torch
from mylibs.kriging import Kriging
from sm_b200 import workloads
X, y = workloads.c5_kriging_r3()
k = Kriging(X, y); k.update_param([0.05, 0.05], [1.9, 1.9])
for _ in range(3):
    out = k.predict_mesh([-2., -2.], [12., 12.], [10000, 10000], return_mse=False, as_device=True)
torch.cuda.synchronize()
print('ok', float(out[12345]))
