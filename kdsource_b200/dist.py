"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for the few
collectives the path needs.

The reference has no distributed backend; every sub-path shards
without a data-path exchange: query points / MLCV test rows / kNN batches /
output-particle ranges are split contiguously over ranks and the source set is
replicated.  The only collectives are a sum over ranks of small result arrays
(MLCV scores: G doubles) or of disjoint slices (bandwidths, densities).
Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU (tests).
"""

import numpy as np


def _td():
    try:
        import torch.distributed as td
    except ImportError:  # pragma: no cover
        return None
    if td.is_available() and td.is_initialized():
        return td
    return None


def rank_world():
    """(rank, world_size) of the default process group, (0, 1) if none."""
    td = _td()
    if td is None:
        return 0, 1
    return td.get_rank(), td.get_world_size()


def shard_range(n, rank, world):
    """Contiguous [lo, hi) share of n units for `rank`; the first n % world
    ranks get one extra unit (same rule as numpy.array_split)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def allreduce_sum_numpy(a):
    """Sum a numpy array over all ranks (result on every rank)."""
    td = _td()
    if td is None or td.get_world_size() == 1:
        return a
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a))
    if td.get_backend() == "nccl":
        t = t.cuda()
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return t.cpu().numpy()


def allgather_rows(a):
    """Concatenate the ranks' arrays along axis 0 (shares may differ in length);
    the result is on every rank."""
    td = _td()
    if td is None or td.get_world_size() == 1:
        return a
    import torch

    a = np.ascontiguousarray(a)
    world = td.get_world_size()
    cuda = td.get_backend() == "nccl"
    n = torch.tensor([a.shape[0]], dtype=torch.int64)
    n = n.cuda() if cuda else n
    sizes = [torch.zeros_like(n) for _ in range(world)]
    td.all_gather(sizes, n)
    sizes = [int(x.item()) for x in sizes]
    pad = np.zeros((max(sizes),) + a.shape[1:], dtype=a.dtype)
    pad[:a.shape[0]] = a
    t = torch.from_numpy(pad)
    t = t.cuda() if cuda else t
    parts = [torch.empty_like(t) for _ in range(world)]
    td.all_gather(parts, t)
    return np.concatenate([p.cpu().numpy()[:k] for p, k in zip(parts, sizes)], axis=0)


def barrier():
    td = _td()
    if td is not None:
        td.barrier()
