"""world_size-2 gloo run of the N>1 host path: contiguous sharding, (nll, index) all_gather with the
reference's first-minimum tie rule (likeli_optimizer.py:79-81), result gather.  The per-rank likelihood
values come from the oracle here (no GPU in this container); on the box the same code runs over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, 'surrogate-models_b200')):
        sys.path.insert(0, p)
    from sm_b200 import dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    td.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'c5_kriging_r3.npz'))
        nll = g['grid_nll']                              # reference values for all 1600 candidates
        lo, hi = dist.shard_range(nll.shape[0], rank, world)
        local = np.where(np.isnan(nll[lo:hi]), np.inf, nll[lo:hi])
        li = int(np.argmin(local))
        lv, lidx = (float(local[li]), lo + li) if np.isfinite(local[li]) else (float('inf'), -1)
        bv, bi = dist.global_first_min(lv, lidx)
        # tie across ranks: both ranks report the same value -> lowest global index wins
        tv, ti = dist.global_first_min(1.25, 100 - rank)
        # a rank with nothing finite never wins
        nv, ni = dist.global_first_min(float('inf') if rank == 0 else 2.0, -1 if rank == 0 else 77)
        gathered = dist.gather_slices(torch.from_numpy(nll[lo:hi].copy()), nll.shape[0])
        ok = np.array_equal(gathered.numpy(), nll, equal_nan=True)
        # the packed 16-byte variant fed from (value, index) tensors: same three answers
        t = lambda v, i: (torch.tensor([v], dtype=torch.float64), torch.tensor([i], dtype=torch.int64))  # noqa: E731
        ok = ok and dist.global_first_min_dev(*t(lv, lidx)) == (bv, bi)
        ok = ok and dist.global_first_min_dev(*t(1.25, 100 - rank)) == (tv, ti)
        ok = ok and dist.global_first_min_dev(*t(float('inf') if rank == 0 else 2.0, -1 if rank == 0 else 77)) == (nv, ni)
        ok = ok and dist.global_first_min_dev(*t(float('nan') if rank == 0 else float('inf'), 5 if rank == 0 else -1)) == (float('inf'), -1)
        ok = ok and dist.global_first_min_dev(*t(-3.5, (1 << 40) + rank)) == (-3.5, 1 << 40)   # indices beyond 2^32 survive
        q.put((rank, bv, bi, tv, ti, nv, ni, ok, dist.rank_world()))
    finally:
        td.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_argmin_and_gather():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bv, bi, tv, ti, nv, ni, ok, rw in res:
        assert bi == 960 and abs(bv - (-111.51605954168531)) < 1e-9
        assert (tv, ti) == (1.25, 99)
        assert (nv, ni) == (2.0, 77)
        assert ok and rw == (rank, 2)
