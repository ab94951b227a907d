#!/bin/bash
# iteration loop: parity tests, the randomised sweep, then the short benchmark
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu.log
timeout 600 python tools/fuzz_parity.py --cases ${FUZZ_CASES:-200} > $O/fuzz.log 2>&1; echo "fuzz rc=$?"; tail -4 $O/fuzz.log
python bench.py --n 200000 --steps 2 --warmup 1 --quick > $O/bench_iter.json 2> $O/bench_iter.err; echo "bench rc=$?"; tail -3 $O/bench_iter.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_iter.json').read().strip().splitlines()[-1])
print("mlcv value", d["value"], "roofline", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
for k, v in d["sub"].items():
    print(k, json.dumps({a: b for a, b in v.items() if a != "roofline"})[:1500], (v.get("roofline") or {}).get("frac"))
PY
