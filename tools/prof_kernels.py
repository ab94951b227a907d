"""Run each hot kernel a few times at a small size (for ncu captures)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_scaled  # noqa: E402
from kdsource_b200 import _lib, kde  # noqa: E402
from kdsource_b200.sampler import DeviceSampler, make_geom_struct  # noqa: E402
from kdsource_b200.treekde import TreeKDE  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
t = _lib.torch()
N = 100_000
parts, w, g, scaling, data = synthetic_scaled(N)
bw = kde.bw_silv(6, np.sum(w) ** 2 / np.sum(w ** 2))
if which in ("all", "eval"):
    M = 200_000
    q = synthetic_scaled(M, seed=54321, wseed=1)[4]
    k = TreeKDE(bw=bw).fit(data, w)  # the default path: ordered tiles, tile-local coordinates
    dev = k._device_state()
    d_q = _lib.to_dev(q)
    for _ in range(3):
        k.evaluate_device(d_q, M, dev)
    t.cuda.synchronize()
if which in ("all", "evalx2"):
    M = 200_000
    q = synthetic_scaled(M, seed=54321, wseed=1)[4]
    k2 = TreeKDE(bw=bw, dtype="float32x2").fit(data, w)
    dev2 = k2._device_state()
    d_q2 = _lib.to_dev(q)
    for _ in range(3):
        k2.evaluate_device(d_q2, M, dev2)
    t.cuda.synchronize()
if which in ("all", "mlcv"):
    plan = kde.MLCVPlan(data, w, bw, np.logspace(-0.3, 0.3, 32), n_splits=10)
    for _ in range(3):
        plan.launch()
    t.cuda.synchronize()
    print(plan.collect()[:4])
if which in ("all", "knn"):
    off = kde.array_split_bounds(N, 10)
    for prec in ("float64", "float32"):
        for _ in range(2):
            kde.knn_batched(data, 11, off, precision=prec)
if which in ("all", "resample"):
    Nr = 10_000_000  # a list that has outgrown the L2 (configuration 4)
    pr, wr, gr, sr, _ = synthetic_scaled(Nr, seed=7)
    bws = np.random.default_rng(3).uniform(0.2, 0.5, Nr).astype(np.float32)
    smp = DeviceSampler(pr, wr, bws, make_geom_struct(gr, sr, "g"))
    for _ in range(3):
        smp.sample_device(100_000_000, seed=1)
    t.cuda.synchronize()
if which in ("all", "gzip"):
    import ctypes

    L = _lib.load()
    bws = np.random.default_rng(3).uniform(0.2, 0.5, N).astype(np.float32)
    smp = DeviceSampler(parts, w, bws, make_geom_struct(g, scaling, "g"))
    n = 10_000_000
    rec = smp.sample_device(n, seed=1)["records"]
    nb, member = n * 36, int(L.kdsb200_gz_member_bytes())
    hist = _lib.empty(258, t.int64)
    _lib.check(L.kdsb200_gz_histogram(_lib.ptr(rec), (64 << 20) // 36 * 36, member, 36, 12, _lib.ptr(hist),
                                      _lib.stream()))
    hh = hist.cpu().numpy().astype(np.uint64)
    table = _lib.GzTable()
    _lib.check(L.kdsb200_gz_build_table(hh.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), 36, 12,
                                        ctypes.byref(table)))
    cap = (nb + nb // 4 + (1 << 20)) // 4 * 4
    out = _lib.empty(cap, t.uint8)
    scr = _lib.empty(int(L.kdsb200_gz_scratch_bytes(nb, member)), t.uint8)
    tot = _lib.empty(1, t.int64)
    for _ in range(3):
        _lib.check(L.kdsb200_gz_compress(_lib.ptr(rec), nb, member, ctypes.byref(table),
                                         _lib.ptr(scr), _lib.ptr(out), cap, _lib.ptr(tot),
                                         _lib.stream()))
    t.cuda.synchronize()
    print("gzip ratio", int(tot.cpu()[0]) / nb)
print("done", which)
