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


def allreduce_sum_tensor(t):
    """In-place sum of a tensor over all ranks: on the tensor's device and current stream
    with NCCL (no host round trip), through the CPU with gloo."""
    td = _td()
    if td is None or td.get_world_size() == 1:
        return t
    if td.get_backend() == "nccl" or not t.is_cuda:
        td.all_reduce(t, op=td.ReduceOp.SUM)
        return t
    c = t.cpu()
    td.all_reduce(c, op=td.ReduceOp.SUM)
    t.copy_(c)
    return t


def allgather_shards(t, n_total):
    """Assemble the full (n_total, ...) tensor on every rank from the ranks' contiguous
    row shares (`shard_range` sizes: they differ by at most one row) -- one all-gather
    of equal-size padded pieces on the tensor's device (NVLink with NCCL)."""
    td = _td()
    if td is None or td.get_world_size() == 1:
        return t
    import torch

    world = td.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    via_cpu = t.is_cuda and td.get_backend() != "nccl"
    src = t.cpu() if via_cpu else t
    pad = torch.zeros((mx,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    pad[:src.shape[0]] = src
    out = torch.empty((world * mx,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    td.all_gather_into_tensor(out, pad)
    full = torch.cat([out[r * mx:r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)
    return full.to(t.device) if via_cpu else full


def allgather_concat_tensor(t):
    """Concatenate the ranks' tensors along dim 0 (shares may differ in length) on the
    tensor's device."""
    td = _td()
    if td is None or td.get_world_size() == 1:
        return t
    import torch

    world = td.get_world_size()
    via_cpu = t.is_cuda and td.get_backend() != "nccl"
    src = t.cpu() if via_cpu else t
    n = torch.tensor([src.shape[0]], dtype=torch.int64, device=src.device)
    ns = torch.empty(world, dtype=torch.int64, device=src.device)
    td.all_gather_into_tensor(ns, n)
    sizes = [int(x) for x in ns.cpu()]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    pad[:src.shape[0]] = src
    out = torch.empty((world * mx,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    td.all_gather_into_tensor(out, pad)
    full = torch.cat([out[r * mx:r * mx + k] for r, k in enumerate(sizes)], dim=0)
    return full.to(t.device) if via_cpu else full


def all_agree(ok):
    """True on every rank iff `ok` is true on every rank (one small all-reduce)."""
    td = _td()
    if td is None or td.get_world_size() == 1:
        return bool(ok)
    return bool(allreduce_sum_numpy(np.array([0.0 if ok else 1.0]))[0] == 0.0)


def barrier():
    td = _td()
    if td is not None:
        td.barrier()
