#!/bin/bash
# Round-end measurement pass on an 8-GPU box: NVLS correctness on real 8-way multicast, device-timed
# sweeps of every collective (CUDA-graph replay), then the bench for both arms (eager, one process per GPU).
set -x
S="python scripts/bw_sweep.py --world 8"
timeout 300 python -m pytest tests/test_gpu_collectives.py -m gpu -q 2>&1 | tail -5
timeout 150 $S --algos ll,oneshot --min 64 --max 65536 --step 4
timeout 150 $S --algos auto --min 1024 --max 1073741824 --step 4
timeout 120 $S --symm --algos auto --min 1048576 --max 1073741824 --step 4
timeout 100 $S --op allgather --min 4096 --max 134217728 --step 8
timeout 100 $S --op reducescatter --min 4096 --max 134217728 --step 8
timeout 100 $S --op broadcast --min 4096 --max 268435456 --step 8
timeout 100 $S --op sendrecv --min 1024 --max 1073741824 --step 8
timeout 100 $S --op grad --min 2097152 --max 134217728 --step 4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --steps 30 --warmup 5 2>gpurun_out/bench_n8.err | tee gpurun_out/bench_n8.json | cut -c1-400
tail -2 gpurun_out/bench_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --impl nccl --gpus 8 --steps 30 --warmup 5 2>gpurun_out/bench_n8_nccl.err | tee gpurun_out/bench_n8_nccl.json | cut -c1-400
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 4 --steps 30 --warmup 5 2>gpurun_out/bench_n4.err | tee gpurun_out/bench_n4.json | cut -c1-400
