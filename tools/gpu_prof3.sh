#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python tools/prof_kernels.py knn > $O/p_knn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 2 -c 2 -o $O/prof_knn2_r01 python tools/prof_kernels.py knn > $O/ncu_knn.log 2>&1; echo "ncu knn rc=$?"
for pm in 0 4 3 38 2; do
  KDSB200_MLCV_POLY=$pm python bench.py --n 200000 --steps 2 --warmup 1 --no-sub > $O/bench_poly_$pm.json 2> $O/bench_poly_$pm.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_poly_$pm.json').read().strip().splitlines()[-1])
print("poly $pm: value", d["value"], "ms", d["ms_per_step"], "argmax", d["scores_argmax"])
PY
done
KDSB200_MLCV_POLY=3 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k mlcv 2>&1 | tail -3
