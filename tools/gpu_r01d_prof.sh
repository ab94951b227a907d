#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
TAG=r01d
python tools/prof_kernels.py all > $O/prof_plain.log 2>&1; echo "plain rc=$?"; tail -2 $O/prof_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_${TAG}.csv python tools/prof_kernels.py all > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for k in gzenc gzsize; do
  case $k in gzenc) pat=gz_encode_kernel;; gzsize) pat=gz_size_kernel;; esac
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -o $O/prof_${k}_${TAG} -f python tools/prof_kernels.py gzip > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
python tools/run_configs.py --configs 4 --scale 10 2>/dev/null | tee $O/configs_c4_full.jsonl | cut -c1-700
python - <<'PY'
import sys, time, os, tempfile
sys.path.insert(0, '.')
import numpy as np, bench
ps, ws_, gs_, sc_, _ = bench.synthetic_scaled(1_000_000, seed=7)
bws = np.random.default_rng(3).uniform(0.2, 0.5, len(ps)).astype(np.float32)
print(bench.resample_e2e(ps, ws_, bws, gs_, sc_, False))
PY
