#!/bin/bash
# One-shot tuning / measurement pass on an 8-GPU box (device-timed, CUDA-graph replayed).
set -x
W=${1:-8}
S="python scripts/bw_sweep.py --world $W"
# A. NVLS reduce phase in isolation (operands in the symmetric heap): CTAs x loads in flight
timeout 120 $S --symm --algos nvls --blocks 64,100,148 --min 67108864 --max 268435456 --step 4 --nvls-unr 4
timeout 120 $S --symm --algos nvls --blocks 64,100,148 --min 67108864 --max 268435456 --step 4 --nvls-unr 8
# B. staged tensors: phase-by-phase vs warp-specialised pipelined kernel
timeout 120 $S --algos nvls --blocks 0 --min 67108864 --max 1073741824 --step 4 --pipe-min 99999999999 --nvls-unr 8
timeout 120 $S --algos nvls --blocks 0 --min 67108864 --max 1073741824 --step 4 --pipe-min 0
# C. the other collectives
timeout 100 $S --op allgather --min 65536 --max 134217728 --step 8
timeout 100 $S --op reducescatter --min 65536 --max 134217728 --step 8
timeout 100 $S --op broadcast --min 65536 --max 268435456 --step 8
timeout 100 $S --op sendrecv --min 1024 --max 1073741824 --step 16
# D. fused gradient kernel (bf16 wire), sizes of the ResNet-50 buckets and beyond
timeout 100 $S --op grad --min 8388608 --max 268435456 --step 4 --pipe-min 99999999999
timeout 100 $S --op grad --min 8388608 --max 268435456 --step 4 --pipe-min 0
