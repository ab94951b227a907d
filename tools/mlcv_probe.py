"""Per-rank MLCV kernel time for the row share of rank r of W (emulated on one GPU) against the
number of source splits."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import D_MLCV, G_MLCV, K_MLCV, synthetic_scaled, time_events  # noqa: E402
from kdsource_b200 import _lib, dist, kde  # noqa: E402

t = _lib.torch()
N = 1_000_000
_, w, _, _, data = synthetic_scaled(N)
bw = kde.bw_silv(D_MLCV, np.sum(w) ** 2 / np.sum(w ** 2))
grid = np.logspace(-0.3, 0.3, G_MLCV)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
splits = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
fb = kde.kfold_bounds(N, K_MLCV)
for r in (0, W // 2):
    lo, hi = dist.shard_range(N, r, W)
    for ns in splits:
        plan = kde.MLCVPlan(data, w, bw, grid, n_splits=K_MLCV, rows=(lo, hi), nsplit=ns or None)
        plan.launch()
        ms = float(np.median(time_events(t, plan.launch, 3)))
        print("rank %d of %d rows [%d, %d): nsplit %d -> %.1f ms, %.3e useful evals/s, %.3e slots/s"
              % (r, W, lo, hi, plan.nsplit, ms, plan.n_evals / ms * 1e3, plan.n_slots / ms * 1e3),
              flush=True)
        del plan
