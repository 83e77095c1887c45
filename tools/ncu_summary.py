#!/usr/bin/env python
"""Print the metrics we track from an .ncu-rep (reads `ncu --page raw --csv`)."""
import csv
import subprocess
import sys

WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__t_bytes.sum l1tex__t_bytes.sum
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__warps_active.avg.pct_of_peak_sustained_active
launch__registers_per_thread launch__grid_size launch__block_size launch__occupancy_limit_shared_mem
launch__occupancy_limit_registers launch__occupancy_limit_warps launch__waves_per_multiprocessor
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__inst_executed_op_shared_atom.sum sm__inst_executed_pipe_lsu.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_membar_per_issue_active.ratio
smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio
smsp__average_warps_issue_stalled_selected_per_issue_active.ratio""".split()


def main(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print("==", name[:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:82s} {vals[i]:>16s} {units[i]}")
        if pat:
            for i, h in enumerate(hdr):
                if pat in h and h not in WANT:
                    print(f"  {h:82s} {vals[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
