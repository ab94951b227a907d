#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/multi_check.py > $O/multi_check.log 2>&1; echo "multi_check rc=$?"; tail -4 $O/multi_check.log
python bench.py --steps 1 --warmup 0 --no-sub > $O/bench_one.json 2> $O/bench_one.err && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mlcv_pairs -c 1 --csv --log-file $O/traffic_mlcv_full.csv python bench.py --steps 1 --warmup 0 --no-sub > $O/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"; tail -4 $O/traffic_mlcv_full.csv
