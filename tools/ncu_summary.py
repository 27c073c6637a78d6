"""Summarise an ncu --page raw --csv dump: python tools/ncu_summary.py raw.csv [pattern ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
data = [r for r in rows[1:] if len(r) == len(hdr)]
units = data[0] if data and not data[0][0].isdigit() else None
kernels = [r for r in data if r[0].isdigit()]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "launch__waves", "achieved_occupancy",
                        "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "pipe_fmaheavy", "pipe_alu.avg.pct",
                        "pipe_xu.avg.pct", "pipe_lsu.avg.pct", "pipe_fp64", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
                        "op_fadd_pred_on.sum", "op_fmul_pred_on.sum", "op_ffma_pred_on.sum", "smsp__inst_executed.sum", "thread_inst_executed_per_inst",
                        "issue_stalled", "local", "shared_mem_per_block", "l1tex__data_bank_conflicts", "lts__t_bytes.sum"]
for k in kernels:
    print("==", k[hdr.index("Kernel Name")][:90], "block", k[hdr.index("Block Size")], "grid", k[hdr.index("Grid Size")])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            v = k[i]
            if v in ("", "n/a"):
                continue
            print(f"  {h} [{units[i] if units else ''}] = {v}")
