#!/bin/bash
# Round-2 evidence for the NVLink kernels at WORLD GPUs (one process drives all ranks):
#  1. scripts/nvlink_evidence.py: device time per launch + NVLink bytes from the driver counters
#  2. ncu --replay-mode app-range around ONE launch of every kernel (range replay does not serialise
#     the kernels inside the range, so co-dependent grids on different GPUs still meet):
#     gpu__time_duration, dram__bytes_read/write, lts__t_bytes, nvlrx/nvltx__bytes of GPU 0
W=${1:-2}
mkdir -p gpurun_out/ncu_n$W
timeout 400 python scripts/nvlink_evidence.py --world $W --out gpurun_out/nvlink_evidence_n$W.json > gpurun_out/nvlink_evidence_n$W.log 2>&1
python - "$W" <<'PY' > gpurun_out/cases_n$W.txt
import json, sys
for r in json.load(open(f"gpurun_out/nvlink_evidence_n{sys.argv[1]}.json")):
    print(r["kernel"])
PY
i=0
while IFS= read -r k; do
  i=$((i+1))
  timeout 200 ncu --replay-mode app-range --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum \
      --csv --log-file "gpurun_out/ncu_n$W/case_$i.csv" python scripts/nvlink_evidence.py --world $W --ncu-range "$k" > gpurun_out/ncu_n$W/case_$i.log 2>&1
  echo "$k" > "gpurun_out/ncu_n$W/case_$i.name"
done < gpurun_out/cases_n$W.txt
ls gpurun_out/ncu_n$W | wc -l
