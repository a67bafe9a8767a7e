"""Analytic test functions used as parity inputs (same definitions as the reference's
myutils/samples.py:18-42; pure host-side fixtures)."""
import math


def f_3d(coords):
    x, y = coords[0], coords[1]
    return math.sin(x) + 0.95 + 0.075*x**2 - 0.001*x**4 + 0.05*y**2 - 0.001*y**4 - 0.005*y**3


def f_2d(x):
    return 6 - (math.sin(x) + 0.95 + 0.075*x**2 - 0.001*x**4)


def f_2d_big(x):
    return math.sin(x)+3.+0.035*x**2+0.003*x**3-0.00045*x**4+0.0000005*x**6


def generate_sample_data(f_3d, px_edge, py_edge):
    px, py, pz = [], [], []
    for ix in range(len(px_edge)):
        for iy in range(len(py_edge)):
            px.append(px_edge[ix])
            py.append(py_edge[iy])
            pz.append(f_3d([px_edge[ix], py_edge[iy]]))
    return px, py, pz
