"""Deterministic synthetic inputs for the BASELINE.json configurations.

The generators restate the reference's own samplers so the inputs are bit-identical to what the
reference scripts use: Halton (mylibs/halton.py:40-112), StructuredSample
(mylibs/structured_sample.py:34-66) and the test functions of myutils/samples.py.  Host-side input
generation only -- no likelihood / prediction arithmetic happens here.
"""
import math

import numpy as np

from myutils.samples import f_2d, f_3d, generate_sample_data


def _primes(count):
    out, num = [2], 3
    while len(out) < count:
        if all(num % q for q in out):
            out.append(num)
        num += 2
    return out[:count]


def halton_point(i, base):
    """Radical inverse of i+1 in `base`, accumulated like mylibs/halton.py:61-71 (least significant digit
    first, digit * base ** -(j+1))."""
    i = int(i) + 1
    res, j = 0., 0
    while i:
        i, digit = divmod(i, base)
        res += int(digit) * int(base) ** (-1 * (j + 1))
        j += 1
    return res


def halton_plan(point_count, dimension, bounds, base=None):
    """Halton().generate_sample_plan (mylibs/halton.py:92-112): bases = first primes."""
    base = list(base) if base is not None else _primes(dimension)
    pts = []
    for i in range(point_count):
        pts.append([bounds[d][0] + (halton_point(i, base[d]) * (bounds[d][1] - bounds[d][0]))
                    for d in range(dimension)])
    return np.array(pts, dtype=np.float64)


def structured_plan(point_count, dimension, bounds):
    """StructuredSample().generate_sample_plan (mylibs/structured_sample.py:34-66), full-factorial."""
    edge = math.ceil(point_count ** (1 / float(dimension)))
    idx, cnt = [], np.zeros(dimension)
    while len(idx) < point_count:
        idx.append([cnt[d] for d in range(dimension)])
        for d in range(dimension):  # little-endian odometer (increase_edge, :22-31)
            if cnt[d] >= edge - 1:
                cnt[d] = 0
            else:
                cnt[d] += 1
                break
    norm = np.array(idx) * (1 / (edge - 1))
    pts = []
    for i in range(point_count):
        row = []
        for d in range(dimension):
            point = norm[i][d]
            if max(norm[:, d]) < 1 and point_count > 2:
                point = point * (1. / max(norm[:, d]))
            row.append(bounds[d][0] + (point * (bounds[d][1] - bounds[d][0])))
        pts.append(row)
    return np.array(pts, dtype=np.float64)


def fb(x):
    """Second C2 response: well-conditioned optimum."""
    return f_3d(x) + 0.5 * math.sin(3 * x[0] + 1) * math.sin(2.3 * x[1])


def c1_kriging_r2():
    """examples/kriging_r2.py:61-69: 8 structured samples of f_2d on [0, 10]."""
    X = structured_plan(8, 1, [(0., 10.)])
    return X, np.array([f_2d(x[0]) for x in X])


def c2_halton64(response='f3d'):
    """n = 64 Halton samples on [0,10]^2, y = f_3d (or `fb`)."""
    X = halton_plan(64, 2, [(0., 10.), (0., 10.)])
    f = f_3d if response == 'f3d' else fb
    return X, np.array([f(x) for x in X])


def c3_lhs1024(n=1024, k=8, seed=0):
    """8-D latin hypercube (scipy qmc; the reference's own LHS allocates n^k), smooth response."""
    from scipy.stats import qmc
    X = qmc.LatinHypercube(d=k, seed=seed).random(n)
    return X, np.sin(3 * X.sum(axis=1)) + X[:, 0] ** 2


def c3_candidates(C=100000, k=8, seed=0):
    """log10 theta ~ U(-3, 1), p ~ U(1, 2): rows of the explicit candidate list."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-3, 1, (C, k)), rng.uniform(1, 2, (C, k))


def c4_rbf4096(n=4096):
    from scipy.stats import qmc
    X = qmc.Halton(d=3, scramble=False).random(n + 1)[1:] * 10
    return X, np.sin(X[:, 0]) + 0.05 * X[:, 1] ** 2 - 0.01 * X[:, 2] ** 3


def c4_queries(Nq, seed=1):
    return np.random.default_rng(seed).uniform(0, 10, (Nq, 3))


def c5_kriging_r3():
    """examples/kriging_r3.py:33-36: 6 x 6 grid on {0,2,..,10}^2, y = f_3d."""
    e = [0., 2., 4., 6., 8., 10.]
    px, py, pz = generate_sample_data(f_3d, e, e)
    return np.array([px, py]).T, np.array(pz)


def c5_query_mesh(m, lo=-2., hi=12., row_begin=0, row_count=None):
    """Rows [row_begin, row_begin + row_count) of the m x m mesh over [lo, hi]^2 (row-major), the plotting
    domain of examples/kriging_r3.py:27-28."""
    row_count = m - row_begin if row_count is None else row_count
    ax = np.linspace(lo, hi, m)
    out = np.empty((row_count, m, 2))
    out[:, :, 0] = ax[row_begin:row_begin + row_count, None]
    out[:, :, 1] = ax[None, :]
    return out.reshape(-1, 2)
