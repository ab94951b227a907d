"""world_size-2 `gloo` tests of the multi-rank host logic (no GPU)."""

import os

import numpy as np
import pytest


def _worker(rank, world, port, q):
    import torch.distributed as td

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from kdsource_b200 import dist

    assert dist.rank_world() == (rank, world)
    lo, hi = dist.shard_range(101, rank, world)
    part = np.zeros(101)
    part[lo:hi] = np.arange(lo, hi)
    tot = dist.allreduce_sum_numpy(part)
    scores = dist.allreduce_sum_numpy(np.array([1.0 + rank, 2.0]))
    dist.barrier()
    q.put((rank, lo, hi, tot.tolist(), scores.tolist()))
    td.destroy_process_group()


def test_shard_and_allreduce_two_ranks():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, tot0, sc0), (r1, lo1, hi1, tot1, sc1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 51, 51, 101)
    assert tot0 == tot1 == list(np.arange(101.0))
    assert sc0 == sc1 == [3.0, 4.0]


def test_shard_range_covers_everything():
    from kdsource_b200.dist import shard_range

    for n in (0, 1, 7, 100, 1000003):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_mlcv_row_sharding_arithmetic():
    """The per-row coefficients used by the sharded MLCV score sum to the
    reference's mean-of-means over folds regardless of how rows are split."""
    from kdsource_b200 import kde
    from kdsource_b200.dist import shard_range

    rng = np.random.default_rng(0)
    N, K = 1003, 10
    w = rng.uniform(0.5, 1.5, N)
    v = rng.normal(size=N)  # stand-in for ln S_i
    fb = kde.kfold_bounds(N, K)
    ref = np.mean([np.mean(w[fb[f]:fb[f + 1]] * v[fb[f]:fb[f + 1]]) for f in range(K)])
    nf = np.diff(fb)
    fold_of = np.repeat(np.arange(K), nf)
    coefw = w / (nf[fold_of] * K)
    for world in (1, 2, 3, 8):
        tot = 0.0
        for r in range(world):
            lo, hi = shard_range(N, r, world)
            tot += np.sum(coefw[lo:hi] * v[lo:hi])
        assert tot == pytest.approx(ref, rel=1e-13)
