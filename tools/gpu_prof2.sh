#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python tools/calib.py > $O/calib2.json 2>&1; cat $O/calib2.json
python tools/prof_kernels.py knn > $O/p_knn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 2 -c 2 -o $O/prof_knn_r01 python tools/prof_kernels.py knn > $O/ncu_knn.log 2>&1; echo "ncu knn rc=$?"
python tools/prof_kernels.py resample > $O/p_rs_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:resample_kernel -s 1 -c 1 -o $O/prof_resample_r01 python tools/prof_kernels.py resample > $O/ncu_rs.log 2>&1; echo "ncu resample rc=$?"
