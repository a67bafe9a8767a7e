"""Multi-GPU plumbing for the two shardable parts of the path: candidates and queries are
independent, so each rank takes a contiguous index range and the only data-path collectives are

  * one all_gather of (nll_min, global_index) per rank -> lexicographic first-minimum on every rank
    (NCCL has no MINLOC; reproduces the strict '<' scan of likeli_optimizer.py:79-81 across shards);
  * an optional all_gather of prediction slices.

Works with any initialised torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as td


def rank_world(distributed=None):
    """(rank, world) -- (0, 1) unless torch.distributed is initialised (or distributed=False)."""
    if distributed is False or not (td.is_available() and td.is_initialized()):
        return 0, 1
    return td.get_rank(), td.get_world_size()


def shard_range(total, rank, world):
    """contiguous range [lo, hi) of `total` items for `rank`; preserves global order for tie-breaking"""
    return total * rank // world, total * (rank + 1) // world


def first_min_pairs(vals, idxs):
    """lexicographic minimum over (value, index) pairs, NaN and index -1 never win; none -> (inf, -1)"""
    best_v, best_i = float('inf'), -1
    for v, i in zip(vals, idxs):
        if i < 0 or v != v:
            continue
        if v < best_v or (v == best_v and (best_i < 0 or i < best_i)):
            best_v, best_i = v, i
    return best_v, best_i


def global_first_min(val, idx, device=None, group=None):
    """all_gather one (val, idx) pair per rank and reduce identically on every rank"""
    if not (td.is_available() and td.is_initialized()) or td.get_world_size(group) == 1:
        return float(val), int(idx)
    world = td.get_world_size(group)
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if td.get_backend(group) == 'nccl' else 'cpu'
    v = torch.tensor([float(val)], dtype=torch.float64, device=device)
    i = torch.tensor([int(idx)], dtype=torch.int64, device=device)
    vs = torch.empty(world, dtype=torch.float64, device=device)
    is_ = torch.empty(world, dtype=torch.int64, device=device)
    td.all_gather_into_tensor(vs, v, group=group)
    td.all_gather_into_tensor(is_, i, group=group)
    return first_min_pairs(vs.tolist(), is_.tolist())


def global_first_min_dev(val_t, idx_t, group=None):
    """Same reduction as global_first_min, fed straight from the kernel's output buffers (val_t: 1 x float64, idx_t:
    1 x int64, on the device for nccl / on the host for gloo): ONE 16-byte all_gather (the index travels as the bit pattern
    of a float64), the lexicographic reduction runs on the gathered tensor, and a single 16-byte copy brings the result
    to the host -- one synchronisation per step instead of the six of the host-scalar version."""
    pair = torch.stack([val_t.reshape(()), idx_t.reshape(()).view(torch.float64)])
    if td.is_available() and td.is_initialized() and td.get_world_size(group) > 1:
        world = td.get_world_size(group)
        out = torch.empty(2 * world, dtype=torch.float64, device=pair.device)
        td.all_gather_into_tensor(out, pair, group=group)
        vals, idxs = out[0::2], out[1::2].contiguous().view(torch.int64)
        inf = torch.full_like(vals, float('inf'))
        vals = torch.where((idxs >= 0) & ~torch.isnan(vals), vals, inf)     # NaN and "nothing finite" never win
        vmin = vals.min()
        imin = torch.where(vals == vmin, idxs, torch.full_like(idxs, torch.iinfo(torch.int64).max)).min()
        imin = torch.where(torch.isinf(vmin), torch.full_like(imin, -1), imin)
        pair = torch.stack([vmin, imin.view(torch.float64)])
    host = pair.cpu()
    return float(host[0]), int(host[1:].view(torch.int64)[0])


def gather_slices(local, total, group=None):
    """concatenate per-rank contiguous slices (shard_range order) of a 1-D result on every rank"""
    if not (td.is_available() and td.is_initialized()) or td.get_world_size(group) == 1:
        return local
    world = td.get_world_size(group)
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    m = max(sizes)
    pad = torch.zeros(m, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    out = torch.empty(world * m, dtype=local.dtype, device=local.device)
    td.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * m:r * m + sizes[r]] for r in range(world)])
