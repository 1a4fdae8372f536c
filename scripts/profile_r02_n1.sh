#!/bin/bash
# Round-2 ncu evidence at N=1 (one GPU; never wrap a multi-rank command in ncu).
#  1. bench line (not under ncu)
#  2. launch list of the bench step with per-launch device time (cold-cache, serialised: compare SHARES)
#  3. --set full capture of grad_local_kernel
#  4. smoke() under ncu (launch list): must stay serialisable
set -x
mkdir -p gpurun_out
export B200_BENCH_WATCHDOG_S=3000
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -1 gpurun_out/bench_n1.json | cut -c1-1500
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_n1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile \
    > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:grad_local_kernel -s 5 -c 5 \
    -o gpurun_out/prof_grad_local python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile \
    > gpurun_out/bench_under_ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_under_ncu.log 2>&1
tail -2 gpurun_out/smoke_under_ncu.log
python -c "import __graft_entry__ as g; g.smoke()"
ls -la gpurun_out/ | head -30
