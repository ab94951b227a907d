#!/bin/bash
# N-GPU evidence run: sharded-path equivalence check, bench.py at N ranks, configuration 3 at full size
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | wc -l; free -g | sed -n 2p; nproc
timeout 600 $TR --master-port 29513 tools/multi_check.py > $O/multi_check_${N}.log 2>&1; echo "multi_check rc=$?"; tail -3 $O/multi_check_${N}.log
timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench rc=$?"; tail -3 $O/bench_${N}gpu.err; cat $O/bench_${N}gpu.json | cut -c1-900
timeout 900 $TR --master-port 29515 tools/run_configs.py --configs 3 > $O/configs_${N}gpu.jsonl 2> $O/configs_${N}gpu.err; echo "configs rc=$?"; tail -3 $O/configs_${N}gpu.err; cat $O/configs_${N}gpu.jsonl
