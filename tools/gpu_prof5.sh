#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python tools/prof_kernels.py knn > $O/p_knn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_small -s 2 -c 1 -o $O/prof_knn4_r01 -f python tools/prof_kernels.py knn > $O/ncu_knn.log 2>&1; echo "ncu knn rc=$?"
