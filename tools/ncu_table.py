"""Prints the metric table of a kernel from an .ncu-rep, or from its `--page raw --csv` export
(developer tool for profiles/*.md).
usage: python tools/ncu_table.py <report.ncu-rep | raw.csv>"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "gpc__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    if sys.argv[1].endswith(".csv"):
        txt = open(sys.argv[1]).read()
    else:
        txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in txt.splitlines() if not l.startswith("==")]))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("## `%s`\n" % vals[hdr.index("Kernel Name")])
    print("| metric | unit | value |\n|---|---|---|")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} | {units[i]} | {vals[i]} |")


if __name__ == "__main__":
    main()
