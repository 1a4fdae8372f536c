"""profiles/r02/nvlink_kernels_n<W>.csv from the artefacts of scripts/profile_r02_nvlink.sh."""
import csv
import glob
import json
import os
import sys

src, dst, W = sys.argv[1], sys.argv[2], sys.argv[3]
ev = {r["kernel"]: r for r in json.load(open(f"{src}/nvlink_evidence_n{W}.json"))}
ncu = {}
for name_file in glob.glob(f"{src}/ncu_n{W}/case_*.name"):
    name = open(name_file).read().strip()
    path = name_file[:-5] + ".csv"
    if not os.path.exists(path):
        continue
    rows = [l for l in open(path) if not l.startswith("==")]
    per_dev = {}
    for r in csv.DictReader(rows):
        if r.get("Metric Name"):
            per_dev.setdefault(r["Device"], {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    if per_dev:
        ncu[name] = per_dev.get("0") or next(iter(per_dev.values()))
cols = ["kernel", "world", "bytes", "us_per_launch(events)", "busbw_GB/s", "nvlink_tx_GB/s(NVML data)", "nvlink_rx_GB/s(NVML data)",
        "nvlink_raw_tx_GB/s", "frac_of_900(raw)", "ncu_range_us", "ncu_dram_read_MB", "ncu_dram_write_MB", "ncu_lts_MB",
        "ncu_nvlrx_MB", "ncu_nvltx_MB", "hbm_GB/s_in_range(of 6482.7)"]
with open(f"{dst}/nvlink_kernels_n{W}.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(cols)
    for k, r in ev.items():
        m = ncu.get(k, {})
        t = m.get("gpu__time_duration.sum")
        rd, wr = m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum")
        raw = max(r.get("nvlink_raw_tx_gbs") or 0, r.get("nvlink_raw_rx_gbs") or 0)
        w.writerow([k, r["world"], r["bytes"], r["us_per_launch"], r["busbw_gbs"], r["nvlink_tx_gbs"], r["nvlink_rx_gbs"],
                    r["nvlink_raw_tx_gbs"], round(raw / 900, 3) if raw else None,
                    round(t / 1e3, 1) if t else None, round(rd / 1e6, 2) if rd is not None else None,
                    round(wr / 1e6, 2) if wr is not None else None,
                    round(m["lts__t_bytes.sum"] / 1e6, 2) if "lts__t_bytes.sum" in m else None,
                    round(m["nvlrx__bytes.sum"] / 1e6, 2) if "nvlrx__bytes.sum" in m else None,
                    round(m["nvltx__bytes.sum"] / 1e6, 2) if "nvltx__bytes.sum" in m else None,
                    round((rd + wr) / t, 1) if t and rd is not None else None])
with open(f"{dst}/ncu_app_range_n{W}.csv", "w") as f:  # the raw per-case ncu rows
    w = csv.writer(f)
    w.writerow(["case", "device", "metric", "unit", "value"])
    for name_file in sorted(glob.glob(f"{src}/ncu_n{W}/case_*.name")):
        path = name_file[:-5] + ".csv"
        if os.path.exists(path):
            for r in csv.DictReader([l for l in open(path) if not l.startswith("==")]):
                if r.get("Metric Name"):
                    w.writerow([open(name_file).read().strip(), r["Device"], r["Metric Name"], r["Metric Unit"], r["Metric Value"]])
print(open(f"{dst}/nvlink_kernels_n{W}.csv").read())
