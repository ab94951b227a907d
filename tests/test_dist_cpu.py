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


def _fit_worker(rank, world, port, q, fn):
    import torch.distributed as td

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    import kdsource_b200.api as kds
    from kdsource_b200 import dist

    rows = dist.allgather_rows(np.full((3 + rank, 2), float(rank)))
    geo = kds.geom.Geometry([kds.geom.Lethargy(10), kds.geom.SurfXY(), kds.geom.Isotrop()])
    s = kds.KDSource(kds.PList(fn, set_params=False), geo, bw=0.4)
    s.fit(N=2500, skip=100, importance=[2, 1, 1, 1, 1, 1], sharded=True)
    q.put((rank, rows.tolist(), s.scaling.tolist(), float(s.N_eff), s.kde.data.tolist(),
           s.kde.weights.tolist()))
    td.destroy_process_group()


def test_sharded_fit_equals_unsharded_fit(tmp_path):
    """KDSource.fit(sharded=True) on 2 ranks (each reads half of the list; N_eff and the
    scaling from all-reduced weighted moments; samples all-gathered) gives the model of
    an unsharded fit, including the pooled std of the Isotrop metric and a list with
    invalid particles in it."""
    import torch.multiprocessing as mp

    import kdsource_b200.api as kds
    from kdsource_b200 import mcpl
    from oracle import kds_oracle as O

    parts, w = O.synthetic_particles(3000, seed=5, wseed=6)
    w[::13] = 0.0  # dropped by PList.get on whichever rank reads them
    fn = str(tmp_path / "list.mcpl.gz")
    mcpl.write_mcpl_file(fn, ekin=parts[:, 0], x=parts[:, 1], y=parts[:, 2], z=parts[:, 3],
                         ux=parts[:, 4], uy=parts[:, 5], uz=parts[:, 6], time=parts[:, 7],
                         weight=w, pdgcode=2112, srcname="t")
    geo = kds.geom.Geometry([kds.geom.Lethargy(10), kds.geom.SurfXY(), kds.geom.Isotrop()])
    ref = kds.KDSource(kds.PList(fn, set_params=False), geo, bw=0.4)
    ref.fit(N=2500, skip=100, importance=[2, 1, 1, 1, 1, 1])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_fit_worker, args=(r, 2, port, q, fn)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rows, scaling, n_eff, data, weights in res:
        assert rows == [[0.0, 0.0]] * 3 + [[1.0, 1.0]] * 4
        assert np.allclose(scaling, ref.scaling, rtol=1e-12)
        assert n_eff == pytest.approx(ref.N_eff, rel=1e-13)
        assert np.array_equal(weights, ref.kde.weights)
        assert np.allclose(data, ref.kde.data, rtol=1e-12, atol=1e-15)
