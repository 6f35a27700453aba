"""Print the judged roofline metrics of every kernel in an .ncu-rep (B200_PROFILING.md metric names)."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "XU (MUFU) %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "ALU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("dram__bytes.sum.per_second", "DRAM B/s"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "TMEM active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math throttle"),
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:150])
    for k, label in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"   {label:22s} {r[i]:>16s} {units[i]}")
