#!/bin/bash
# end-of-round refresh: GPU parity suite, smoke, full bench (both arms), launch list of bench.py
cd "$(dirname "$0")/.."
O=gpurun_out
TAG=${1:-r01e}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_${TAG}.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_${TAG}.log
( time python bench.py > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 $O/bench_${TAG}.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_${TAG}.json 2> $O/bench_ref_${TAG}.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-sub > $O/ncu_bench_${TAG}.log 2>&1; echo "ncu bench launches rc=$?"
python tools/run_configs.py --configs 1,2 2>/dev/null | tee $O/configs_12_${TAG}.jsonl | cut -c1-500
