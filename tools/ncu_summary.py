#!/usr/bin/env python
"""Text summary of an ncu report (`ncu --set full --import-source on ...`): the metrics DESIGN.md quotes for every
captured launch, and the source lines with the most warp-stall samples.  Needs `ncu` (reads the report, no GPU).
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, io, subprocess, sys
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, top=16):
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    head, units, launches = rows[0], rows[1], rows[2:]
    kn = head.index("Kernel Name")
    for li, r in enumerate(launches):
        print(f"== launch {li}: {r[kn][:110]}")
        for m in METRICS:
            if m in head:
                i = head.index(m)
                print(f"  {m:86s} {r[i]:>12s} {units[i]}")
        src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                                               "--launch-skip", str(li), "--launch-count", "1"]))))
        per_line = {}
        for s, row in enumerate(src):
            if "# Samples" in row and row[0] == "Line No":
                si = row.index("# Samples")
                for r2 in src[s + 1:]:
                    if len(r2) <= si or r2 and r2[0] == "Line No":
                        break
                    try:
                        n = int(r2[si])
                    except ValueError:
                        continue
                    if r2[0]:
                        per_line[(r2[0], r2[1].strip()[:110])] = per_line.get((r2[0], r2[1].strip()[:110]), 0) + n
        tot = sum(per_line.values()) or 1
        print(f"== hottest source lines (warp stall samples, total {tot})")
        for (ln, text), n in sorted(per_line.items(), key=lambda kv: -kv[1])[:top]:
            print(f"  {100.0 * n / tot:5.1f}%  line {ln}: {text}")


if __name__ == "__main__":
    main(sys.argv[1])
