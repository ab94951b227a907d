#!/bin/bash
# ncu evidence: launch list + full captures of the hot kernels (tools/prof_kernels.py), DRAM
# traffic of the headline kernel at full size.  Usage: gpurun -- bash tools/gpu_prof.sh r02
cd "$(dirname "$0")/.."
O=gpurun_out
TAG=${1:-r02}
python tools/prof_kernels.py all > $O/prof_plain.log 2>&1; echo "plain rc=$?"; tail -2 $O/prof_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_prof_${TAG}.csv python tools/prof_kernels.py all > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for k in ${KERNELS:-mlcv eval knn resample gzenc gzsize}; do
  case $k in
    mlcv) pat=mlcv_pairs_kernel; arg=mlcv;;
    eval) pat=kde_eval_sorted; arg=eval;;
    knn) pat=knn_; arg=knn;;
    resample) pat=resample_kernel; arg=resample;;
    gzenc) pat=gz_encode_kernel; arg=gzip;;
    gzsize) pat=gz_size_kernel; arg=gzip;;
  esac
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -o $O/prof_${k}_${TAG} -f python tools/prof_kernels.py $arg > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
  # the reports are large (gpurun_out/ is capped at 64 MiB): export the pages, drop the report
  R=$O/prof_${k}_${TAG}.ncu-rep
  ncu -i $R --page details > $O/${TAG}_ncu_${k}_details.txt 2>/dev/null
  ncu -i $R --page raw --csv > $O/${TAG}_ncu_${k}_raw.csv 2>/dev/null
  ncu -i $R --page source --csv > $O/${TAG}_ncu_${k}_source.csv 2>/dev/null
  rm -f $R
done
# DRAM traffic of the headline kernel at full size (roofline.traffic in bench.py reads profiles/traffic_mlcv.json)
if [ -z "$SKIP_TRAFFIC" ]; then
python bench.py --steps 1 --warmup 0 --no-sub > $O/bench_one.json 2> $O/bench_one.err && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mlcv_ -c 2 --csv --log-file $O/traffic_mlcv_full.csv python bench.py --steps 1 --warmup 0 --no-sub > $O/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"; tail -8 $O/traffic_mlcv_full.csv | cut -c1-300
fi
