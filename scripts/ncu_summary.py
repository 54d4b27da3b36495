"""Subset of an `ncu --set full` capture as a small CSV for profiles/: python scripts/ncu_summary.py rep [row] > out.csv"""
import csv
import subprocess
import sys

rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.split("\n")))
h, units, r = rows[0], rows[1], rows[2 + row]
keep = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
print(f"# kernel: {r[h.index('Kernel Name')]}")
print("metric,value,unit")
for k in keep:
    if k in h:
        print(f"{k},{r[h.index(k)].replace(',', '')},{units[h.index(k)]}")
for i, k in enumerate(h):
    if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
        print(f"{k},{r[i].replace(',', '')},samples")
