#!/bin/bash
# ncu evidence for the N=1 bench (one GPU; never wrap a multi-rank command in ncu).
#  1. launch list with per-launch device time (cold-cache, serialised: compare SHARES)
#  2. one --set full capture of this repository's dominant kernel at N=1
set -x
mkdir -p gpurun_out
export B200_BENCH_WATCHDOG_S=3000
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_n1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile \
    > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:grad_local_kernel -s 5 -c 3 \
    -o gpurun_out/prof_grad_local python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile \
    > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out/
