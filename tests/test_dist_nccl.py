"""2-rank NCCL run (two B200s of one box, `gpurun --gpus 2 -- python -m pytest tests/test_dist_nccl.py -m gpu`) of the sharded
product API: list-mode likelihood (north-star config 3: candidates split over the ranks), sharded prediction (config 5) and
RBF evaluation, each compared with the single-GPU result of the same call -- bit for bit, the shards are independent."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, 'surrogate-models_b200')):
        sys.path.insert(0, p)
    import torch.distributed as td
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    td.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from mylibs.kriging import Kriging
        from mylibs.rbf import RBF
        from sm_b200 import workloads
        X, y = workloads.c3_lhs1024(n=200, k=4)
        e, p = workloads.c3_candidates(1001, 4)                       # odd count: ragged shards
        e[17, 2] = np.nan                                              # NaN parameters -> NaN, never the minimum
        k = Kriging(X, y)
        full, bv1, bi1 = k.likelihood_batch(e, p, return_best=True, distributed=False)
        shard, bv2, bi2 = k.likelihood_batch(e, p, return_best=True)
        mine = k.likelihood_batch(e, p, gather=False)
        lo, hi = 1001 * rank // world, 1001 * (rank + 1) // world
        ok = {}
        ok['list nll'] = np.array_equal(full, shard, equal_nan=True) and bool(np.isnan(shard[17]))
        ok['list argmin'] = (bv1, bi1) == (bv2, bi2)
        ok['list local slice'] = np.array_equal(mine, full[lo:hi], equal_nan=True)
        param, idx = k.grid_search(4, 2)                               # 4^4 * 2^4 = 4096 candidates over 2 ranks
        param1, idx1 = k.grid_search(4, 2, distributed=False)
        ok['grid'] = idx == idx1 and param == param1
        k.update_param(10. ** e[3], p[3])
        xq = np.random.default_rng(5).uniform(0, 1, (3001, 4))
        y1, m1 = k.predict_batch(xq, return_mse=True, distributed=False)
        y2, m2 = k.predict_batch(xq, return_mse=True)
        ok['predict mse'] = np.array_equal(y1, y2) and np.array_equal(m1, m2)
        y3 = k.predict_batch(xq, distributed=False)
        ok['predict mean slice'] = np.array_equal(k.predict_batch(xq, gather=False), y3[3001 * rank // world:3001 * (rank + 1) // world])
        ok['predict mean'] = np.array_equal(k.predict_batch(xq), y3)
        r = RBF(X, y)
        r.update_param(1.0, 'gaus')
        ok['rbf'] = np.array_equal(r.predict_batch(xq), r.predict_batch(xq, distributed=False))
        bad = [name for name, v in ok.items() if not v]
        q.put((rank, bad, int(bi2)))
    finally:
        td.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sharded_likelihood_and_prediction():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    import torch.multiprocessing as mp
    world, port = 2, 29600 + os.getpid() % 2000
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(not bad for _, bad, _ in res), res
    assert res[0][2] == res[1][2]
