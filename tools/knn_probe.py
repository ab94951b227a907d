"""kNN kernel rates at the bench shape (N=1e6, k=10, batches of 1e4) -- GPU box only."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from kdsource_b200 import _lib, kde  # noqa: E402

t = _lib.torch()
L = _lib.load()
Nk = 1_000_000
dk = bench.synthetic_scaled(Nk, seed=99)[4]
off = kde.array_split_bounds(Nk, 100)
d_data = _lib.to_dev(dk)
d_off = _lib.to_dev(off, np.int64)
out = {"tile": os.environ.get("KDSB200_KNN_TILE", "default")}
for prec, use64 in (("f64", 1), ("f32", 0)):
    scratch = _lib.empty(Nk * 6 + 8 * 200, t.float64 if use64 else t.float32)
    d_kth = _lib.empty(Nk, t.float64)

    def kn():
        _lib.check(L.kdsb200_knn(_lib.ptr(d_data), Nk, 6, _lib.ptr(d_off), len(off) - 1,
                                 int(np.max(np.diff(off))), 11, use64, _lib.ptr(scratch),
                                 _lib.ptr(d_kth), None, None, _lib.stream()))

    for _ in range(2):
        kn()
    ms = float(np.median(bench.time_events(t, kn, 5)))
    out[prec] = {"pairs_per_s": 1e10 / (ms * 1e-3), "ms": ms}
print(json.dumps(out))
