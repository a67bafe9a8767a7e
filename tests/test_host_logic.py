"""CPU-side checks: input generators are bit-identical to the reference's, candidate enumeration of the
host mirror matches the oracle, the opaque-callable path of LikeliOptimizer reproduces the reference's
grid guess, the C-ABI library loads and exports every symbol of include/*.h, and the product refuses to
compute without a CUDA device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import numpy_oracle as orc
from sm_b200 import _cabi, dist, workloads
from mylibs.kriging import Kriging, grid_axes, grid_param
from mylibs.likeli_optimizer import LikeliOptimizer
from mylibs.rbf import RBF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workloads_bit_identical_to_reference_inputs(gold):
    g = gold('c1_kriging_r2')
    X, y = workloads.c1_kriging_r2()
    assert np.array_equal(X, g['X']) and np.array_equal(y, g['y'])
    g = gold('c2_halton64_default_grid')
    X, y = workloads.c2_halton64('f3d')
    assert np.array_equal(X, g['X']) and np.array_equal(y, g['y_f3d'])
    assert np.array_equal(workloads.c2_halton64('fb')[1], g['y_fb'])
    g = gold('c5_kriging_r3')
    X, y = workloads.c5_kriging_r3()
    assert np.array_equal(X, g['X']) and np.array_equal(y, g['y'])
    g = gold('c1_halton8')
    assert np.array_equal(workloads.halton_plan(8, 1, [(0., 10.)]), g['X'])
    g = gold('c3_lhs1024')
    X, y = workloads.c3_lhs1024()
    assert float(X.sum()) == float(g['X_checksum']) and float(y.sum()) == float(g['y_checksum'])
    e, q = workloads.c3_candidates()
    assert np.array_equal(e[g['sel']], g['e_sel']) and np.array_equal(q[g['sel']], g['p_sel'])
    g = gold('c4_rbf4096')
    X, y = workloads.c4_rbf4096()
    assert float(X.sum()) == float(g['X_checksum']) and float(y.sum()) == float(g['y_checksum'])


def test_query_mesh_rows():
    full = workloads.c5_query_mesh(7)
    assert full.shape == (49, 2)
    part = workloads.c5_query_mesh(7, row_begin=2, row_count=3)
    assert np.array_equal(part, full[14:35])


def test_grid_param_matches_oracle_enumeration():
    for k, T, P in ((1, 8, 5), (2, 8, 5), (2, 40, 25), (3, 4, 3)):
        e, q = orc.grid_candidates(k, T, P)
        total = e.shape[0]
        for c in (0, 1, T, total // 3, total - 1):
            assert grid_param(k, T, P, c) == list(e[c]) + list(q[c])
    th, p = grid_axes(8, 5)
    assert np.array_equal(th, np.linspace(-5, 5, 8)) and np.array_equal(p, np.linspace(1., 2., 5))


def test_likeli_optimizer_opaque_callable_matches_reference(gold):
    """func is an arbitrary callable -> the reference's own loop (likeli_optimizer.py:59-86)."""
    g = gold('c1_kriging_r2')
    X, y = g['X'], g['y']
    calls = []

    def func(param):
        calls.append(list(param))
        return orc.nll_opti_exp(X, y, np.array(param))

    opt = LikeliOptimizer()
    guess = opt.generate_grid(func, 1, 8, 5)
    assert len(calls) == 40
    assert np.array_equal(np.array(guess), g['grid_guess'])
    e, q = orc.grid_candidates(1, 8, 5)
    assert np.array_equal(np.array(calls), np.concatenate([e, q], axis=1))   # evaluation order
    assert opt.pBounds == (1., 2.) and opt.thetaBoundsExp == (-5, 3) and opt.maxIter == 5e3
    # all-inf grid -> None (likeli_optimizer.py:51-52)
    assert opt.generate_grid(lambda prm: float('inf'), 1, 3, 2) is None
    # ties: first strict minimum wins
    assert opt.generate_grid(lambda prm: 1.0, 1, 3, 2) == [-5.0, 1.0]
    assert list(opt.increase_i(np.array([7., 0.]), 8)) == [0., 1.]
    assert list(opt.increase_i(np.array([7., 7.]), 8)) == [0., 0.]


def test_likeli_optimizer_find_on_opaque_callable(gold):
    g = gold('c1_kriging_r2')
    X, y = g['X'], g['y']
    res = LikeliOptimizer().find(lambda prm: orc.nll_opti_exp(X, y, np.array(prm)), 1)
    # in-repo known answer (examples/kriging_r2.py:77-79), SLSQP agrees to ~2e-5 across SciPy versions
    assert abs(10. ** res.x[0] - 0.056389969498335794) < 1e-4
    assert abs(res.x[1] - 1.999966138140631) < 1e-4


def test_cabi_exports_every_header_symbol():
    header = open(os.path.join(ROOT, 'include', 'surrogate_b200.h')).read()
    declared = set(re.findall(r'\b(sm_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    L = _cabi.lib()                                   # loads without a GPU
    for name in declared:
        assert hasattr(L, name), name
    assert L.sm_b200_version() >= 100
    assert L.sm_krig_npad(36) == 40 and L.sm_krig_mop_doubles(36) == 15 * 64
    assert b'bad argument' in L.sm_b200_error_string(-1)
    # argument validation happens before any CUDA call
    assert L.sm_krig_nll_grid(None, None, 0, 0, None, 0, None, 0, 0, 0, None, None, None, None) == -1
    assert L.sm_rbf_predict(None, 1, 1, 9, 1.0, None, None, 0, None, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback():
    X, y = workloads.c5_kriging_r3()
    k = Kriging(X, y)                      # construction is host-only, like the reference
    assert k._n == 36 and k._k == 2 and list(k._theta) == [1., 1.] and list(k._p) == [2., 2.]
    assert k._cor_mat is None and k._core_mat_inv is None and k._mu is None
    with pytest.raises(_cabi.SurrogateB200Error):
        k.update_param([1., 1.], [2., 2.])
    with pytest.raises(_cabi.SurrogateB200Error):
        k.grid_search()
    r = RBF(X, y)
    assert r.get_coeff() is None and r.train() is None
    with pytest.raises(_cabi.SurrogateB200Error):
        r.update_param(1.0, 'gaus')
    with pytest.raises(Exception):
        Kriging(X, y).optimize(opti_algo='nonsense')


def test_one_dimensional_inputs_reshaped():
    k = Kriging([0., 1., 2.], [1., 2., 3.])
    assert k._known_in.shape == (3, 1) and k._k == 1
    r = RBF([0., 1., 2.], [1., 2., 3.])
    assert r._known_in.shape == (3, 1)


def test_shard_range_and_first_min_pairs():
    for total in (0, 1, 7, 1600, 10 ** 6):
        for world in (1, 2, 3, 8):
            parts = [dist.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert dist.first_min_pairs([3., 1., 1., float('nan')], [5, 9, 7, 0]) == (1., 7)
    assert dist.first_min_pairs([float('inf'), float('inf')], [-1, -1]) == (float('inf'), -1)
    assert dist.rank_world() == (0, 1)


def test_sampler_mirrors_bit_identical(gold):
    from mylibs.halton import Halton
    from mylibs.structured_sample import StructuredSample
    g = gold('validation_r3')
    assert np.array_equal(np.array(Halton().generate_sample_plan(7, 3, [(0., 1.), (-1., 1.), (2., 5.)])), g['halton_7x3'])
    assert np.array_equal(np.array(StructuredSample().generate_sample_plan(9, 2, [(0., 10.), (-5., 5.)])), g['structured_9x2'])
    assert Halton().prime(0) == 2 and Halton().prime(3) == 7


def test_validation_with_opaque_callable(gold):
    """Validation on an arbitrary callable follows the reference's per-point path"""
    from mylibs.validation import Validation, ValidationResults
    g = gold('validation_r3')
    f = lambda x: float(np.sin(x[0]) + 0.05 * x[1] ** 2)   # noqa: E731
    v = Validation()
    vx, vy = g['vx'], g['vy']
    pred = np.array([f(x) for x in vx])
    assert abs(v.calc_rmse(vx, vy, f) - np.sqrt(np.mean((vy - pred) ** 2))) < 1e-14
    assert abs(v.calc_mae(vx, vy, f) - np.max(np.abs(vy - pred))) < 1e-14
    assert abs(v.calc_deviation(vx, vy, f) - np.mean(np.abs(vy - pred)) / vy.mean()) < 1e-14
    assert np.allclose(v.calc_rae(vx, vy, f), np.abs((vy - pred) / vy), rtol=0, atol=1e-14)
    assert ValidationResults().press == 0.


def test_exp2_core_emulation(tmp_path):
    """csrc/fastmath2.cuh (base-2 table-driven power / exponential of the n > 144 likelihood kernel) emulated on the
    host with the same constants and operation order (tools/exp2_check.c, C99 fma) against long-double libm"""
    import json
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / 'exp2_check')
    subprocess.run(['gcc', '-O2', '-std=gnu99', '-ffp-contract=off', os.path.join(root, 'tools', 'exp2_check.c'), '-lm', '-o', exe],
                   check=True)
    out = json.loads(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    assert out['pow2_fast_max_ulp'] < 1.1 and out['exp2_neg_max_ulp'] < 1.1
    assert out['zero_dist'] < 1e-100 and out['padding'] > 1e200          # |dx| = 0 -> 0, padding -> "infinity"
    assert out['exp2_neg_0'] == 1.0 and out['exp2_neg_1100'] == 0.0 and out['exp2_neg_inf'] == 0.0
    # the kernel's constants are the ones the emulation uses
    hdr = open(os.path.join(root, 'surrogate-models_b200', 'csrc', 'fastmath2.cuh')).read()
    src = open(os.path.join(root, 'tools', 'exp2_check.c')).read()
    for c in ('1.3538030870311431e-03', '9.163913992275266e-07', '4.135378350676714e-10', '1.399621994296973e-13'):
        assert c in hdr and c in src


def test_odometer_and_timer_semantics():
    """increase_i: little-endian, in place, wraps to zeros (reference likeli_optimizer.py:89-97); TimeTrack interface
    (reference myutils/time_track.py:18-38)."""
    import time
    from mylibs.likeli_optimizer import LikeliOptimizer
    from myutils.time_track import TimeTrack
    lo = LikeliOptimizer()
    d = np.zeros(3)
    seen = []
    for _ in range(27):
        seen.append(tuple(int(v) for v in d))
        r = lo.increase_i(d, 3)
        assert r is d
    assert len(set(seen)) == 27 and seen[1] == (1, 0, 0) and seen[3] == (0, 1, 0) and seen[9] == (0, 0, 1)
    assert tuple(d) == (0, 0, 0)
    assert (lo.pBounds, lo.thetaBoundsExp, lo.maxIter, lo.multi_start) == ((1., 2.), (-5, 3), 5e3, 1)
    t = TimeTrack('unit')
    time.sleep(0.01)
    assert t.get_time() >= 0.01
    t.tic()
    lap = t.toc()
    assert 0 <= lap < 0.01 and t.get_time() >= lap


def test_cta_kernel_slot_packing_replay(tmp_path):
    """csrc/cta_slots.h (live-range packing of the factor tiles in krig_nll_cta.cu) compiled for the host: a replay of the kernel's
    read / store schedule (tools/cta_slots_check.cpp) never finds a tile overwritten before its last read, for every tile count the
    kernel accepts and beyond; the slot count is NT^2/4 + O(NT) (71 at n = 128, 271 at n = 256)"""
    import ctypes
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / 'libcta_slots.so')
    subprocess.run(['g++', '-O2', '-shared', '-fPIC', os.path.join(root, 'tools', 'cta_slots_check.cpp'), '-o', so], check=True)
    lib = ctypes.CDLL(so)
    out = (ctypes.c_int * 40)()
    assert lib.cta_slots_check(2, 41, out) == 0
    slots = dict(zip(range(2, 42), out))
    assert slots[16] == 71 and slots[32] == 271 and slots[22] == 131
    for nt in range(4, 42):
        assert slots[nt] <= nt * nt // 4 + nt and slots[nt] < nt * (nt - 1) // 2 + 1
