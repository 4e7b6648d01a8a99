#!/usr/bin/env python
"""Text summary of an `ncu --set full --import-source on` report: headline metrics + hottest source lines.
Usage: python profiles/summarize.py gpurun_out/<name>.ncu-rep > profiles/<name>.txt   (needs ncu on PATH)"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    for k, row in enumerate(rows[2:]):
        name = row[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        print(f"== launch {k}: {name[:110]}")
        for w in WANT:
            if w in h:
                i = h.index(w)
                print(f"  {w:80s} {row[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    cur, out = None, []
    for r in csv.reader(io.StringIO(src)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 5 and r[0].isdigit() and r[4].replace(",", "").isdigit():
            out.append((int(r[4].replace(",", "")), cur, int(r[0]), r[1].strip()[:110]))
    tot = sum(o[0] for o in out) or 1
    print(f"== hottest source lines (warp stall samples, total {tot})")
    for s, f, l, text in sorted(out, reverse=True)[:25]:
        print(f"  {100 * s / tot:5.1f}%  {f}:{l}  {text}")


if __name__ == "__main__":
    main(sys.argv[1])
