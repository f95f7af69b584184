#!/usr/bin/env python
"""Key metrics + hottest instructions of one .ncu-rep (read on the CPU box: `ncu -i ... --page raw/source --csv`)."""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_xu.sum"]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units, row = r[0], r[1], r[2]
    print("kernel:", row[hdr.index("Kernel Name")][:110])
    for i, h in enumerate(hdr):
        if h in KEEP:
            print(f"  {h} = {row[i]} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h2 = rows[1]
    data = rows[2:]
    iS, iI, iSrc = h2.index("# Samples"), h2.index("Instructions Executed"), h2.index("Source")
    tot = sum(int(d[iS]) for d in data) or 1
    print(f"  total samples {tot}, total warp-instructions {sum(int(d[iI]) for d in data)}")
    for k in sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:top]:
        print(f"  {100.0 * int(data[k][iS]) / tot:5.1f}%  inst={data[k][iI]:>10}  {data[k][iSrc][:100]}")


if __name__ == "__main__":
    main()
