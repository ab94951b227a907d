#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "2gpu rc=$?"; tail -c 1200 $O/bench_2gpu.json; tail -5 $O/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/bench_2gpu_ref.json 2> $O/bench_2gpu_ref.err; echo "2gpu ref rc=$?"; cat $O/bench_2gpu_ref.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/multi_check.py > $O/multi_check.log 2>&1; echo "multi_check rc=$?"; tail -5 $O/multi_check.log
