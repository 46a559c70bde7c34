#!/usr/bin/env python
"""Per-kernel table of the metrics this repo's profiles quote, from an .ncu-rep with several kernels:
    python tools/ncu_summary.py rep.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__pcsamp_sample_count",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "sm__cycles_active.min", "sm__cycles_active.max",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [r[ix["Kernel Name"]].split("(")[0][-40:] for r in body]
    print("%-62s %-10s %s" % ("metric", "unit", " | ".join(names)))
    for m in METRICS:
        if m not in ix:
            continue
        print("%-62s %-10s %s" % (m, units[ix[m]], " | ".join(r[ix[m]] for r in body)))


if __name__ == "__main__":
    main()
