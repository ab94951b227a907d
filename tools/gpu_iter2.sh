#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
for q in 4 2; do
KDSB200_EVAL_Q=$q python bench.py --n 100000 --steps 1 --warmup 1 > $O/bench_q$q.json 2> $O/bench_q$q.err; echo "bench q=$q rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_q$q.json').read().strip().splitlines()[-1])
for k, v in d["sub"].items():
    print("Q=$q", k, {a: b for a, b in v.items() if a not in ("roofline","workload")}, (v.get("roofline") or {}).get("frac"))
PY
done
