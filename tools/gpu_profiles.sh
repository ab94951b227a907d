#!/bin/bash
# Full benchmark + the ncu evidence committed under profiles/ (one GPU)
cd "$(dirname "$0")/.."
O=gpurun_out
TAG=${1:-r01}
python bench.py > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err; echo "bench rc=$?"; tail -3 $O/bench_${TAG}.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_${TAG}.json 2> $O/bench_ref_${TAG}.err; echo "ref rc=$?"
python tools/prof_kernels.py all > $O/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_${TAG}.csv python tools/prof_kernels.py all > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for k in mlcv eval knn resample; do
  case $k in mlcv) pat=mlcv_pairs;; eval) pat=kde_eval_kernel;; knn) pat=knn_small;; resample) pat=resample_kernel;; esac
  python tools/prof_kernels.py $k > $O/prof_plain_$k.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -o $O/prof_${k}_${TAG} -f python tools/prof_kernels.py $k > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
