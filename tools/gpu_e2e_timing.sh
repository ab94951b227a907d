#!/bin/bash
cd "$(dirname "$0")/.."
KDSB200_TIMING=1 python - <<'PY' 2>&1 | grep -E "timing|value"
import sys, time, os, tempfile, contextlib
sys.path.insert(0, '.')
import numpy as np, bench
ps, ws_, gs_, sc_, _ = bench.synthetic_scaled(1_000_000, seed=7)
bws = np.random.default_rng(3).uniform(0.2, 0.5, len(ps)).astype(np.float32)
from kdsource_b200 import _lib
_lib.torch().zeros(10, device="cuda")
for i in range(2):
    print("--- run", i)
    print(bench.resample_e2e(ps, ws_, bws, gs_, sc_, False))
PY
