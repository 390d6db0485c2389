"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.
usage: python scripts/summarise_profiles.py <tag>   (reads gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_*.ncu-rep)"""
import collections
import csv
import glob
import os
import subprocess
import sys

tag = sys.argv[1]
os.makedirs("profiles", exist_ok=True)
# ---- launch list ----
src = f"gpurun_out/{tag}_launches.csv"
if os.path.exists(src):
    lines = [l for l in open(src) if not l.startswith("==")]
    open(f"profiles/{tag}_launches.csv", "w").writelines(lines)
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0]); a[0] += 1; a[1] += v
    with open(f"profiles/{tag}_launches_summary.md", "w") as f:
        f.write(f"# {tag} - ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n"
                "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv python bench.py --steps 4 --warmup 3 "
                "--equil 3 --fcm-steps 4 --no-cpu-baseline`\n(cold-cache, serialised launches: compare SHARES, not absolutes). "
                f"Raw CSV: profiles/{tag}_launches.csv\n\n| kernel | launches | avg us | total us |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k[:110]}` | {n} | {t / n:.1f} | {t:.1f} |\n")
# ---- full captures: the metrics the design cites ----
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
for rep in sorted(glob.glob(f"gpurun_out/{tag}_*.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, data = rows[0], rows[1], rows[2:]
    name = os.path.basename(rep).replace(".ncu-rep", "")
    with open(f"profiles/{name}_raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        cols = [hdr.index("Kernel Name")] + [hdr.index(m) for m in WANT if m in hdr]
        w.writerow([hdr[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for d in data:
            w.writerow([d[c] for c in cols])
    print("wrote", f"profiles/{name}_raw.csv", len(data), "kernels")
