"""Summarise an ncu report (.ncu-rep) as a markdown table of the metrics the roofline discussion uses.

    python scripts/ncu_summary.py [--all] gpurun_out/X.ncu-rep "title" > profiles/X.md
"""
import csv
import subprocess
import sys

METRICS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    argv = [a for a in sys.argv[1:] if a != "--all"]
    every = "--all" in sys.argv  # every captured launch instead of the first of each (kernel, grid)
    rep, title = argv[0], (argv[1] if len(argv) > 1 else argv[0])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keep, seen = [], {}
    for i, r in enumerate(data):  # first launch of every distinct (kernel, grid)
        key = (r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")])
        if every or key not in seen:
            seen[key] = i
            keep.append(i)
    print(f"# {title}\n")
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in keep) + " |")
    print("|---|---|" + "---|" * len(keep))
    for m in METRICS:
        if m not in hdr:
            continue
        j = hdr.index(m)
        vals = [data[i][j] for i in keep]
        if m == "Kernel Name":
            vals = [v.replace("void tsfmx::<unnamed>::", "").replace("void unnamed>::", "")[:48] for v in vals]
        print(f"| `{m}` | {units[j]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
