#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "knn" 2>&1 | tail -2
python - <<'PY'
import numpy as np, sys, time
sys.path.insert(0, '.')
from bench import synthetic_scaled, time_events
from kdsource_b200 import _lib, kde
L = _lib.load(); t = _lib.torch()
for Nk in (1_000_000,):
    _, wk, _, _, dk = synthetic_scaled(Nk, seed=99)
    off = kde.array_split_bounds(Nk, int(round(Nk / 10000)))
    d_data = _lib.to_dev(dk); d_off = _lib.to_dev(off, np.int64)
    for prec, use64 in (("f64", 1), ("f32", 0)):
        scratch = _lib.empty(Nk * 6, t.float64 if use64 else t.float32)
        d_kth = _lib.empty(Nk, t.float64)
        def kn():
            _lib.check(L.kdsb200_knn(_lib.ptr(d_data), Nk, 6, _lib.ptr(d_off), len(off) - 1, int(np.max(np.diff(off))), 11, use64, _lib.ptr(scratch), _lib.ptr(d_kth), None, None, _lib.stream()))
        for _ in range(2): kn()
        ms = float(np.median(time_events(t, kn, 5)))
        pk = float(np.sum(np.diff(off).astype(np.float64) ** 2))
        print(Nk, prec, "pairs/s %.4g" % (pk / (ms * 1e-3)), "ms %.3f" % ms)
PY
