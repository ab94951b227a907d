#!/bin/bash
# Round-1 final refresh: GPU parity suite, smoke, full bench (both arms), ncu launch list of
# bench.py, launch list + full captures of the hot kernels (incl. the gzip kernels)
cd "$(dirname "$0")/.."
O=gpurun_out
TAG=${1:-r01d}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_${TAG}.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_${TAG}.log
( time python bench.py > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 $O/bench_${TAG}.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_${TAG}.json 2> $O/bench_ref_${TAG}.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-sub > $O/ncu_bench_${TAG}.log 2>&1; echo "ncu bench launches rc=$?"
python tools/prof_kernels.py all > $O/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_${TAG}.csv python tools/prof_kernels.py all > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for k in gzenc gzsize; do
  case $k in gzenc) pat=gz_encode_kernel;; gzsize) pat=gz_size_kernel;; esac
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -o $O/prof_${k}_${TAG} -f python tools/prof_kernels.py gzip > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
tail -c 400 $O/bench_${TAG}.json
