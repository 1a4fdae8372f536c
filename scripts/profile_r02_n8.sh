#!/bin/bash
# 8-GPU evidence, kept short (8x the GPU-minutes): NVML NVLink counters for every kernel, ncu
# app-range for the three kernels that only exist with >= 3 ranks and the multicast mapping.
W=${1:-8}
mkdir -p gpurun_out/ncu_n$W
timeout 300 python scripts/nvlink_evidence.py --world $W --out gpurun_out/nvlink_evidence_n$W.json > gpurun_out/nvlink_evidence_n$W.log 2>&1
tail -3 gpurun_out/nvlink_evidence_n$W.log | cut -c1-200
i=0
for k in "allreduce_pipe_kernel<NVLS> 256MiB" "allreduce_twoshot_kernel<NVLS> 64MiB" "allgather_pull_kernel 64MiB total" "grad_allreduce_kernel"; do
  i=$((i+1))
  timeout 150 ncu --replay-mode app-range --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,nvlrx__bytes.sum,nvltx__bytes.sum \
      --csv --log-file "gpurun_out/ncu_n$W/case_$i.csv" python scripts/nvlink_evidence.py --world $W --ncu-range "$k" > gpurun_out/ncu_n$W/case_$i.log 2>&1
  python - "$W" "$k" "$i" <<'PY'
import json, sys
W, k, i = sys.argv[1:]
names = [r["kernel"] for r in json.load(open(f"gpurun_out/nvlink_evidence_n{W}.json"))]
full = [n for n in names if k in n]
open(f"gpurun_out/ncu_n{W}/case_{i}.name", "w").write(full[0] if full else k)
PY
done
ls gpurun_out/ncu_n$W
