"""Turns the ncu artefacts of scripts/profile_n1.sh into the tracked summaries under profiles/.

    python scripts/summarize_profile.py gpurun_out profiles/r01
"""
import collections
import csv
import re
import subprocess
import sys


def main(src, dst):
    with open(f"{src}/launches_n1.csv") as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    ours = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ns = v * 1e3 if unit in ("usecond", "us") else (v * 1e6 if unit in ("msecond", "ms") else v)
        short = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
        tot[short] += ns
        cnt[short] += 1
        if "b200::" in row["Kernel Name"]:
            ours.append((row["ID"], short, row["Grid Size"], row["Block Size"], ns / 1e3))
    total = sum(tot.values())
    with open(f"{dst}/launches_n1_summary.md", "w") as out:
        out.write("# ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile` (N=1)\n\n")
        out.write("`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and "
                  "serialised, so compare shares, not absolutes.\n\n")
        out.write(f"{sum(cnt.values())} launches, {total / 1e6:.2f} ms of kernel time in total.\n\n")
        out.write("| share | total ms | launches | kernel |\n|---|---|---|---|\n")
        for k, v in tot.most_common(15):
            out.write(f"| {100 * v / total:.1f}% | {v / 1e6:.3f} | {cnt[k]} | `{k}` |\n")
        out.write("\n## This repository's kernels\n\n| launch id | kernel | grid | block | us |\n|---|---|---|---|---|\n")
        for o in ours:
            out.write(f"| {o[0]} | `{o[1]}` | {o[2]} | {o[3]} | {o[4]:.2f} |\n")
        mine = sum(o[4] for o in ours) * 1e3
        out.write(f"\n{len(ours)} launches, {mine / 1e6:.3f} ms = {100 * mine / total:.2f}% of the kernel time.\n")
    raw = subprocess.run(["ncu", "-i", f"{src}/prof_grad_local.ncu-rep", "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(f"{dst}/grad_local_kernel_ncu_full.csv", "w") as out:
        w = csv.writer(out)
        w.writerow([hdr[i] for i in idx])
        w.writerow([rows[1][i] for i in idx])  # units
        for row in rows[2:]:
            w.writerow([row[i] for i in idx])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
