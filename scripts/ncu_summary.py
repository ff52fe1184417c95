"""Key metrics of every launch in an ncu report (`ncu --set full ... -o X`), one column per launch (kept under profiles/).

usage: python scripts/ncu_summary.py gpurun_out/X.ncu-rep [title]
"""
import csv, io, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ik = hdr.index("Kernel Name")
if len(sys.argv) > 2:
    print("#", sys.argv[2])
print("# launches:", " | ".join(r[ik].replace("madipm::", "").split("(")[0] for r in data))
extra = [h for h in hdr if ("tensor" in h and "pct_of_peak_sustained_active" in h and h not in METRICS)]
for m in METRICS + extra:
    if m not in hdr:
        continue
    i = hdr.index(m)
    print(f"{m + ' [' + units[i] + ']':100s} " + " ".join(f"{r[i]:>14s}" for r in data))
