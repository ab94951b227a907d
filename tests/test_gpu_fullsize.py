"""BASELINE.json-size runs checked through size-independent properties and
oracle spot checks (the oracle cannot run these sizes in full)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scaled(n, seed=12345, wseed=777):
    from bench import synthetic_scaled

    return synthetic_scaled(n, seed=seed, wseed=wseed)


def test_config1_verification_example(oracle):
    """configs[0]: 1e5-particle list, Silverman bw, evaluate on a 1e4 grid."""
    from kdsource_b200 import kde
    from kdsource_b200.treekde import TreeKDE

    _, w, _, _, data = _scaled(100_000)
    q = _scaled(10_000, seed=54321, wseed=1)[4]
    bw = kde.bw_silv(6, np.sum(w) ** 2 / np.sum(w ** 2))
    got = TreeKDE(bw=bw).fit(data, w).evaluate(q)
    ref = oracle.kde_sum_c(data, w, bw, q[:1500])
    ok = ref > 1e-12 * ref.max()
    assert np.max(np.abs(got[:1500][ok] / ref[ok] - 1)) < 1e-5


def test_evaluate_1e6_sources_linearity_and_spot_check(oracle):
    from kdsource_b200.treekde import TreeKDE

    _, w, _, _, data = _scaled(1_000_000)
    q = _scaled(200_000, seed=54321, wseed=1)[4]
    bw = np.random.default_rng(0).uniform(0.15, 0.3, len(data))
    full = TreeKDE(bw=bw).fit(data, w).evaluate(q)
    h = len(data) // 2
    a = TreeKDE(bw=bw[:h]).fit(data[:h], w[:h]).evaluate(q)
    b = TreeKDE(bw=bw[h:]).fit(data[h:], w[h:]).evaluate(q)
    wa, wb = w[:h].sum() / w.sum(), w[h:].sum() / w.sum()
    # three independent fp32 evaluations (different recentring): 1e-5 relative
    mix = wa * a + wb * b
    ok = mix > 1e-12 * mix.max()
    rel = np.abs(full[ok] / mix[ok] - 1)
    assert rel.max() < 1e-5, rel.max()
    pick = np.random.default_rng(1).choice(len(q), 48, replace=False)
    ref = oracle.kde_sum_c(data, w, bw, q[pick])
    assert np.max(np.abs(full[pick] / ref - 1)) < 1e-5


def test_config2_mlcv_1e6_row_contributions(oracle):
    """configs[1] (N=1e6, d=6, G=32, K=10): the contribution of a block of test
    rows against the oracle, and additivity of disjoint row blocks."""
    from kdsource_b200 import kde

    _, w, _, _, data = _scaled(1_000_000)
    N = len(data)
    bw = kde.bw_silv(6, np.sum(w) ** 2 / np.sum(w ** 2))
    grid = np.logspace(-0.3, 0.3, 32)
    r0 = 299_990  # straddles the fold boundary at 300000
    rows = (r0, r0 + 40)
    plan = kde.MLCVPlan(data, w, bw, grid, n_splits=10, rows=rows)
    plan.launch()
    got = plan.collect()
    fb = oracle.kfold_bounds(N, 10)
    ref = np.zeros(len(grid))
    idx = np.arange(*rows)
    fold_of = np.searchsorted(fb, idx, side="right") - 1
    for f in np.unique(fold_of):
        mine = idx[fold_of == f]
        train = np.r_[0:fb[f], fb[f + 1]:N]
        dtr, wtr = data[train], w[train]
        for g, gv in enumerate(grid):
            dens = oracle.kde_sum_c(dtr, wtr, gv * bw, data[mine])
            ref[g] += np.sum(w[mine] * np.log(dens)) / ((fb[f + 1] - fb[f]) * 10)
    assert np.allclose(got, ref, rtol=1e-5)
    # additivity over row blocks (what the multi-GPU sharding relies on)
    blocks = [(0, 2000), (2000, 3000), (3000, 4096)]
    parts = []
    for blk in blocks:
        p = kde.MLCVPlan(data, w, bw, grid, n_splits=10, rows=blk)
        p.launch()
        parts.append(p.collect())
    whole = kde.MLCVPlan(data, w, bw, grid, n_splits=10, rows=(0, 4096))
    whole.launch()
    assert np.allclose(whole.collect(), np.sum(parts, axis=0), rtol=1e-10)


def test_config3_knn_1e7(oracle):
    """configs[2], first half: kNN adaptive bandwidth, k=10, on a 1e7-particle
    list in the reference's default batches of 1e4."""
    from kdsource_b200 import kde

    N = 10_000_000
    rng = np.random.default_rng(5)
    data = rng.normal(size=(N, 6))
    data[:, 0] += np.where(rng.uniform(size=N) < 0.5, 2.0, -2.0)
    bw = kde.bw_knn(data, np.ones(N), k=10, batch_size=10000)
    assert bw.shape == (N,) and np.all(bw > 0) and np.all(np.isfinite(bw))
    off = kde.array_split_bounds(N, 1000)
    for b in (0, 417, 999):
        s, e = off[b], off[b + 1]
        rd, _ = oracle.knn(data[s:e], 11)
        assert np.array_equal(bw[s:e], rd[:, -1])  # bit-exact against the fp64 oracle


def test_config4_resample_1e7_sources(oracle):
    """configs[3]: adaptive-bandwidth source with 1e7 particles, 2e8 new particles
    in chunks; checks on the device, index parity on a slice."""
    import torch

    from kdsource_b200.sampler import DeviceSampler, make_geom_struct

    parts, w, g, scaling, _ = _scaled(10_000_000, seed=7)
    w = w * np.where(np.arange(len(w)) % 1000 == 0, 50.0, 1.0)  # heavy entries
    bw = np.random.default_rng(3).uniform(0.2, 0.5, len(parts)).astype(np.float32)
    smp = DeviceSampler(parts, w, bw, make_geom_struct(g, scaling, "g"))
    n = 100_000_000
    counts = torch.zeros(1000, dtype=torch.int64, device="cuda")
    keep = None
    for c in range(2):
        r = smp.sample_device(n, seed=11, first=c * n, want_idx=True)
        rec = r["records"].view(torch.int32).view(n, 9)
        assert bool((rec[:, 8] == 2112).all())
        assert bool((rec[:, 7] == 0x3F800000).all())  # weight 1.0f
        assert bool((rec[:, 2] == 0).all()) and bool((rec[:, 6] == 0).all())  # z, time
        counts += torch.bincount(r["idx"] % 1000, minlength=1000)
        if c == 0:
            keep = rec[n - 1000:].clone()
    # heavy entries (index % 1000 == 0) carry 50/(999+50) of the weight
    frac = counts[0].item() / counts.sum().item()
    expect = w[::1000].sum() / w.sum()
    assert frac == pytest.approx(expect, rel=2e-3)
    # chunk boundaries do not matter: a straddling call reproduces both sides
    r = smp.sample_device(2000, seed=11, first=n - 1000, want_idx=True)
    rec = r["records"].view(torch.int32).view(2000, 9)
    assert bool((rec[:1000] == keep).all())
    # index parity with the oracle on a slice deep into the stream
    first = 123_456_789
    got = smp.sample(5000, seed=11, first=first, want=("idx",))["idx"]
    cdf = oracle.weights_cdf(w)
    ref = np.empty(5000, dtype=np.int64)
    tot = int(cdf[-1])
    for j in range(5000):
        p = first + j
        wd = oracle.philox4x32_10([p & 0xFFFFFFFF, p >> 32, 0, 0], [11, 0])
        T = (((wd[0] << 32) | wd[1]) * tot) >> 64
        ref[j] = np.searchsorted(cdf, np.uint64(T), side="right")
    assert np.array_equal(got, ref)
