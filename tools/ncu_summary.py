#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers DESIGN.md cites."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "local_load", "smsp__inst_executed_op_local"]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s  (id %s)" % (d.get("Kernel Name", "?")[:90], d.get("ID")))
        for k, u in zip(hdr, units):
            if any(k == kk or (kk in k and "." not in kk) for kk in KEYS):
                v = d[k]
                if v not in ("", "0", "0.000000"):
                    print("   %-95s %s %s" % (k, v, u))
if __name__ == "__main__":
    main(sys.argv[1])
