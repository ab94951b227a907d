#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py --n 200000 --steps 2 --warmup 1 --quick > $O/bench_quick.json 2> $O/bench_quick.err; echo "quick rc=$?"; tail -c 1500 $O/bench_quick.json; tail -5 $O/bench_quick.err
python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "full rc=$?"; tail -c 3000 $O/bench_full.json; tail -5 $O/bench_full.err
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cat $O/bench_ref.json
python tools/prof_kernels.py all > $O/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r01.csv python tools/prof_kernels.py all > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python tools/prof_kernels.py mlcv > $O/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mlcv_pairs -s 1 -c 1 -o $O/prof_mlcv_r01 python tools/prof_kernels.py mlcv > $O/ncu_mlcv.log 2>&1; echo "ncu mlcv rc=$?"
python tools/prof_kernels.py eval > $O/prof_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:kde_eval_kernel -s 1 -c 1 -o $O/prof_eval_r01 python tools/prof_kernels.py eval > $O/ncu_eval.log 2>&1; echo "ncu eval rc=$?"
ls -la $O
