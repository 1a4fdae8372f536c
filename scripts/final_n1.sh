#!/bin/bash
set -x
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -q -m gpu) > gpurun_out/n1_pytest_final.log 2>&1
tail -4 gpurun_out/n1_pytest_final.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:grad_local_kernel -s 5 -c 5 \
    -o gpurun_out/prof_grad_local python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile > gpurun_out/bench_under_ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
