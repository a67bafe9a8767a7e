"""BASELINE config 5 on one GPU: 10^8 mesh queries, value + MSE -- arbitrary-point path (mesh generated on the device) vs tabulated mesh path"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'surrogate-models_b200'))
import numpy as np, torch
from mylibs.kriging import Kriging
from sm_b200 import device, workloads
X, y = workloads.c5_kriging_r3()
k = Kriging(X, y); k.update_param([0.05, 0.05], [1.9, 1.9])
M = 10000
def generic():
    xs = device.mesh_points([-2., -2.], [12., 12.], [M, M]); return k.predict_device(xs, True)
def mesh():
    return k.predict_mesh([-2., -2.], [12., 12.], [M, M], as_device=True)
def generic_mean():
    xs = device.mesh_points([-2., -2.], [12., 12.], [M, M]); return k.predict_device(xs, False)
def mesh_mean():
    return k.predict_mesh([-2., -2.], [12., 12.], [M, M], return_mse=False, as_device=True)
for name, fn in (('arbitrary-point kernel + device mesh', generic), ('tabulated mesh path', mesh), ('mean only, arbitrary-point kernel + device mesh', generic_mean), ('mean only, tabulated mesh path', mesh_mean)):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); del out
    print(f'{name}: {min(ts):.2f} ms for 1e8 queries = {1e8 / min(ts) * 1e3:.4g} predictions/s')
