"""Randomised parity sweep of the CUDA path against the CPU oracle (GPU box only).

    python tools/fuzz_parity.py [--cases 120] [--seed 0]

Draws random shapes (dimension 1..8, source / query counts around tile boundaries),
weights (including zeros), scalar and per-particle bandwidths, fold counts and grids,
k and batch layouts, metric combinations and kernels, and compares evaluate (fp32 and
fp64), MLCV scores, kNN (bit-exact) and the resampler (fp64 bit-level indices, 1e-9
coordinates; fp32 1e-4 of scale) with the oracle.  Prints the worst deviations and
exits non-zero on the first violation of the tolerances used in tests/.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kdsource_b200 import geom as kgeom  # noqa: E402
from kdsource_b200 import kde  # noqa: E402
from kdsource_b200.sampler import DeviceSampler, make_geom_struct  # noqa: E402
from kdsource_b200.treekde import TreeKDE  # noqa: E402
from oracle import kds_oracle as O  # noqa: E402


def rand_n(rng, lo, hi):
    base = int(rng.integers(lo, hi))
    if rng.uniform() < 0.4:  # near a tile boundary
        base = int(rng.choice([256, 512, 1024, 2048, 4096])) + int(rng.integers(-2, 3))
    return max(1, base)


def case_evaluate(rng, worst):
    D = int(rng.integers(1, 9))
    N, M = rand_n(rng, 1, 6000), rand_n(rng, 1, 3000)
    data = rng.normal(size=(N, D)) * rng.uniform(0.5, 2)
    q = rng.normal(size=(M, D)) * rng.uniform(0.5, 2)
    w = rng.uniform(0.1, 2, N) if rng.uniform() < 0.7 else None
    if w is not None and N > 3 and rng.uniform() < 0.3:
        w[rng.integers(0, N, N // 3)] = 0.0
        w[0] = 1.0
    bw = rng.uniform(0.3, 1.2) if rng.uniform() < 0.5 else rng.uniform(0.3, 1.2, N)
    kern = str(rng.choice(["gaussian", "gaussian", "epa", "box"]))
    ref = O.kde_sum_c(data, w, bw, q) if kern == "gaussian" else O.kde_sum(data, w, bw, q, kernel=kern)
    g64 = TreeKDE(kernel=kern, bw=bw, dtype="float64").fit(data, w).evaluate(q)
    ok = ref > 1e-250
    e64 = float(np.max(np.abs(g64[ok] / ref[ok] - 1))) if ok.any() else 0.0
    assert e64 < 1e-11, ("evaluate fp64", kern, D, N, M, e64)
    worst["evaluate_fp64"] = max(worst.get("evaluate_fp64", 0), e64)
    if kern != "box":
        g32 = TreeKDE(kernel=kern, bw=bw).fit(data, w).evaluate(q)
        big = ref > (1e-12 if kern == "gaussian" else 1e-2) * ref.max()
        e32 = float(np.max(np.abs(g32[big] / ref[big] - 1))) if big.any() else 0.0
        # 1e-5 is the Gaussian tolerance of north_star; 1 - r^2/R^2 cancels near the support
        # radius, so in fp32 the Epanechnikov sum is held to 2e-4 (use dtype="float64" beyond)
        assert e32 < (1e-5 if kern == "gaussian" else 2e-4), ("evaluate fp32", kern, D, N, M, e32)
        worst["evaluate_fp32_" + kern] = max(worst.get("evaluate_fp32_" + kern, 0), e32)
    if kern == "gaussian":
        gx2 = TreeKDE(bw=bw, dtype="float32x2").fit(data, w).evaluate(q)
        big = ref > 1e-12 * ref.max()
        ex2 = float(np.max(np.abs(gx2[big] / ref[big] - 1))) if big.any() else 0.0
        # what remains is the fp32 rounding of the exponent itself: ~1e-7 * |exponent|, i.e. a
        # few 1e-6 for far-tail densities (1e-12 of the maximum)
        assert ex2 < 1e-5, ("evaluate float32x2", D, N, M, ex2)
        worst["evaluate_fp32x2"] = max(worst.get("evaluate_fp32x2", 0), ex2)


def case_mlcv(rng, worst):
    D = int(rng.integers(1, 9))
    N = rand_n(rng, 12, 2500)
    data = rng.normal(size=(N, D))
    w = rng.uniform(0.3, 2, N) if rng.uniform() < 0.7 else None
    seed = rng.uniform(0.5, 1.0) if rng.uniform() < 0.5 else rng.uniform(0.5, 1.0, N)
    G = int(rng.choice([1, 3, 8, 10, 16, 32, 37]))
    grid = np.sort(rng.uniform(0.7, 1.6, G))
    K = int(rng.choice([2, 5, 10, N]))
    got = kde.mlcv_scores(data, w, seed, grid, n_splits=K)
    ref = O.mlcv_scores_c(data, w, seed, grid, min(K, N))
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin), ("mlcv finite pattern", D, N, G, K)
    err = float(np.max(np.abs(got[fin] - ref[fin]) / np.maximum(1.0, np.abs(ref[fin])))) if fin.any() else 0
    assert err < 1e-5, ("mlcv", D, N, G, K, err)
    worst["mlcv"] = max(worst.get("mlcv", 0), err)


def case_knn(rng, worst):
    D = int(rng.integers(1, 9))
    N = rand_n(rng, 40, 5000)
    data = rng.normal(size=(N, D))
    if rng.uniform() < 0.3:
        data = np.round(data * 2) / 2  # exact ties
    nb = int(rng.integers(1, max(2, N // 40)))
    off = kde.array_split_bounds(N, nb)
    kk = int(rng.integers(1, min(40, int(np.min(np.diff(off)))) + 1))
    kth, dist, idx = kde.knn_batched(data, kk, off, want_dist=True, want_idx=True)
    rd, ri = O.knn(data, kk, off)
    assert np.array_equal(dist, rd) and np.array_equal(idx, ri), ("knn", D, N, nb, kk)
    k32 = kde.knn_batched(data, kk, off, precision="float32")[0]
    # fast mode: coordinates are rounded to fp32 (2^-24 relative to |x| ~ 1..4), which bounds
    # the absolute error of a distance; 1e-5 relative once distances are not tiny
    e = float(np.max(np.abs(k32 - rd[:, -1]) / (rd[:, -1] + 0.1)))
    assert e < 1e-5, ("knn fp32", D, N, kk, e)
    worst["knn_fp32"] = max(worst.get("knn_fp32", 0), e)


METRIC_SETS = [
    lambda: ([kgeom.Lethargy(10), kgeom.SurfXY(), kgeom.Isotrop(keep_zdir=True)],
             [("Lethargy", [0.7], [10.0]), ("SurfXY", [3.0, 2.0], [-np.inf, np.inf, -np.inf, np.inf, 0]),
              ("Isotrop", [0.2] * 3, [0, 0, 1])]),
    lambda: ([kgeom.Energy(), kgeom.Vol(-5, 5, -3, 3, -1, 1), kgeom.Isotrop()],
             [("Energy", [0.3], []), ("Vol", [1.0, 2.0, 0.5], [-5, 5, -3, 3, -1, 1]),
              ("Isotrop", [0.2] * 3, [0, 0, 0])]),
    lambda: ([kgeom.Wavelength(), kgeom.SurfR(0, 8), kgeom.Polar(), kgeom.Decade()],
             [("Wavelength", [0.2], []), ("SurfR", [1.0, 20.0], [0, 8, -np.pi, np.pi, 0]),
              ("Polar", [5.0, 10.0], []), ("Decade", [0.3], [])]),
    lambda: ([kgeom.Lethargy(10), kgeom.SurfR2(0, 8), kgeom.PolarMu(), kgeom.Time()],
             [("Lethargy", [0.7], [10.0]), ("SurfR2", [5.0, 20.0], [0, 8, -np.pi, np.pi, 0]),
              ("PolarMu", [0.05, 20.0], []), ("Time", [0.5], [])]),
    lambda: ([kgeom.Lethargy(10), kgeom.SurfCircle(0, 8), kgeom.Isotrop(keep_zdir=True)],
             [("Lethargy", [0.7], [10.0]), ("SurfCircle", [1.0, 1.0], [0, 8, -np.pi, np.pi, 0]),
              ("Isotrop", [np.inf] * 3, [0, 0, 1])]),
]


def case_resample(rng, worst):
    n = rand_n(rng, 1, 4000)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    parts = np.column_stack((rng.uniform(1e-3, 5, n), rng.uniform(-4, 4, n), rng.uniform(-2, 2, n),
                             rng.uniform(-0.9, 0.9, n), d, rng.uniform(0.1, 5, n)))
    w = rng.uniform(0.01, 3, n) ** 2
    metrics, desc = METRIC_SETS[int(rng.integers(0, len(METRIC_SETS)))]()
    kernel = str(rng.choice(["g", "g", "e", "b"]))
    kw = {}
    if rng.uniform() < 0.5:
        kw = dict(trasl=[0.3, -0.2, 0.1], rot=[0.1, -0.2, 0.05])
    g = kgeom.Geometry(metrics, **kw)
    scaling = np.concatenate([np.asarray(s, dtype=np.float64) for _, s, _ in desc])
    bw = rng.uniform(0.2, 0.8) if rng.uniform() < 0.5 else rng.uniform(0.2, 0.8, n).astype(np.float32)
    gs = make_geom_struct(g, scaling, kernel)
    og = O.make_geom(desc, kernel=kernel, **kw)
    count, seed, first = int(rng.integers(1, 6000)), int(rng.integers(0, 2 ** 62)), int(rng.integers(0, 2 ** 40))
    ref = O.resample(parts, w, bw, og, count, seed=seed, first=first)
    o64 = DeviceSampler(parts, w, bw, gs, precision="float64").sample(count, seed=seed, first=first,
                                                                      want=("idx", "parts"))
    assert np.array_equal(o64["idx"], ref["idx"]), ("resample idx fp64", n, count)
    e64 = float(np.max(np.abs(o64["parts"] - ref["parts"]) / (1 + np.abs(ref["parts"]))))
    assert e64 < 1e-9, ("resample parts fp64", desc[1][0], kernel, e64)
    o32 = DeviceSampler(parts, w, bw, gs).sample(count, seed=seed, first=first, want=("idx", "parts"))
    assert np.array_equal(o32["idx"], ref["idx"]), ("resample idx fp32", n, count)
    # precision="float32": every coordinate of every particle within 1e-5 of the variable's
    # scale (fp32 arithmetic for the two specialised geometries, the fp64 store otherwise)
    dev = np.abs(o32["parts"] - ref["parts"]) / (1 + np.abs(ref["parts"]))
    e32 = float(np.max(dev))
    assert e32 < 1e-5, ("resample parts fp32", desc[1][0], kernel, e32,
                        np.unravel_index(np.argmax(dev), dev.shape))
    worst["resample_fp64"] = max(worst.get("resample_fp64", 0), e64)
    worst["resample_fp32"] = max(worst.get("resample_fp32", 0), e32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=120)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    worst = {}
    fns = [case_evaluate, case_mlcv, case_knn, case_resample]
    for i in range(args.cases):
        fns[i % len(fns)](rng, worst)
    print("fuzz ok:", args.cases, "cases; worst deviations:", {k: float("%.3g" % v) for k, v in worst.items()})


if __name__ == "__main__":
    main()
