#!/bin/bash
# N-GPU evidence run: sharded-path equivalence check, bench.py at N ranks, configuration 5
# (or the configurations in $CONFIGS) at full size.  Usage: gpurun --gpus 8 -- bash tools/gpu_multi.sh 8 r02d
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-8}
TAG=${2:-r02}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | wc -l; free -g | sed -n 2p; nproc
timeout 240 $TR --master-port 29513 tools/multi_check.py > $O/multi_check_${N}_${TAG}.log 2>&1; echo "multi_check rc=$?"; tail -3 $O/multi_check_${N}_${TAG}.log
timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_${N}gpu_${TAG}.json 2> $O/bench_${N}gpu_${TAG}.err; echo "bench rc=$?"; tail -3 $O/bench_${N}gpu_${TAG}.err; cat $O/bench_${N}gpu_${TAG}.json | cut -c1-900
timeout ${C_TIMEOUT:-480} $TR --master-port 29515 tools/run_configs.py --configs ${CONFIGS:-5} ${C_ARGS} > $O/configs_${N}gpu_${TAG}.jsonl 2> $O/configs_${N}gpu_${TAG}.err; echo "configs rc=$?"; tail -3 $O/configs_${N}gpu_${TAG}.err; cat $O/configs_${N}gpu_${TAG}.jsonl
