#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "knn" > $O/pytest_knn.log 2>&1; echo "pytest knn rc=$?"; tail -3 $O/pytest_knn.log
python tools/prof_kernels.py resample > $O/p_rs_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:resample_kernel -s 1 -c 1 -o $O/prof_resample2_r01 python tools/prof_kernels.py resample > $O/ncu_rs.log 2>&1; echo "ncu resample rc=$?"
python tools/prof_kernels.py knn > $O/p_knn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_small -s 2 -c 2 -o $O/prof_knn3_r01 python tools/prof_kernels.py knn > $O/ncu_knn.log 2>&1; echo "ncu knn rc=$?"
python bench.py --n 200000 --steps 2 --warmup 1 > $O/bench_iter.json 2> $O/bench_iter.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_iter.json').read().strip().splitlines()[-1])
for k, v in d["sub"].items():
    print(k, {a: b for a, b in v.items() if a != "roofline"}, (v.get("roofline") or {}).get("frac"))
PY
