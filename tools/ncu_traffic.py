"""Extract per-launch DRAM traffic of each profiled kernel from an ncu --set full report into profiles/ncu_traffic.json.
usage: python tools/ncu_traffic.py report.ncu-rep envs dtype"""
import csv
import json
import os
import subprocess
import sys

rep, envs, dtype = sys.argv[1], int(sys.argv[2]), sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
try:
    tab = json.load(open(out_path))
except (OSError, ValueError):
    tab = {"kernels": []}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def packed_thread_instructions(index):
    """Predicated-on thread instructions of the packed fp32 opcodes (FADD2 / FMUL2 / FFMA2, two results each) from the SASS page:
    the op_fadd / op_fmul / op_ffma hardware counters count the scalar opcodes only (checked: counter x cycles = the scalar
    opcode's thread count exactly)."""
    src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    starts = [i for i, q in enumerate(src) if q and q[0] == "Kernel Name"] + [len(src)]
    out = {"FADD2": 0, "FMUL2": 0, "FFMA2": 0}
    for a, b in list(zip(starts[:-1], starts[1:]))[index:index + 1]:  # the source page lists the launches in the raw page's order
        h = src[a + 1]
        si, ti = h.index("Source"), h.index("Predicated-On Thread Instructions Executed")
        for q in src[a + 2:b]:
            if len(q) != len(h) or not q[si].split():
                continue
            t = q[si].split()
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            if op in out:
                out[op] += int(q[ti] or 0)
        break
    return out


launch_index = -1
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    launch_index += 1
    name = r[hdr.index("Kernel Name")]
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(m)
        tot += float(r[i].replace(",", "")) * scale[units[i]]
    dur = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
    row = {"kernel": name.split("(")[0].replace("void ", ""), "envs": envs, "dtype": dtype, "dram_bytes_per_launch": tot,
           "ncu_duration_" + units[hdr.index("gpu__time_duration.sum")]: dur, "report": os.path.basename(rep)}
    # the hardware's own FP fraction: executed (add + mul + 2 fma) thread-instructions per cycle / (2 x fma peak per cycle)
    val = lambda m: float(r[hdr.index(m)].replace(",", "")) if m in hdr else None
    p = "d" if dtype == "f64" else "f"
    rates = [val(f"smsp__sass_thread_inst_executed_op_{p}{op}_pred_on.sum.per_cycle_elapsed") for op in ("add", "mul", "fma")]
    peak = val(f"sm__sass_thread_inst_executed_op_{p}fma_pred_on.sum.peak_sustained")
    if None not in rates and peak:
        cyc = val("sm__cycles_elapsed.avg")
        pk = packed_thread_instructions(launch_index) if (dtype != "f64" and cyc) else {"FADD2": 0, "FMUL2": 0, "FFMA2": 0}
        p2 = [pk[k] / cyc if cyc else 0.0 for k in ("FADD2", "FMUL2", "FFMA2")]
        row.update({"fp_add_per_cycle": rates[0], "fp_mul_per_cycle": rates[1], "fp_fma_per_cycle": rates[2], "fp_fma_peak_per_cycle": peak,
                    "fp_add2_per_cycle": p2[0], "fp_mul2_per_cycle": p2[1], "fp_fma2_per_cycle": p2[2],
                    "fp_frac_counters_scalar_only": (rates[0] + rates[1] + 2 * rates[2]) / (2 * peak),
                    "fp_frac_counters": (rates[0] + rates[1] + 2 * rates[2] + 2 * (p2[0] + p2[1]) + 4 * p2[2]) / (2 * peak)})
    for m, key in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
                   ("launch__registers_per_thread", "registers_per_thread"), ("smsp__inst_executed.sum", "warp_instructions")):
        if val(m) is not None:
            row[key] = val(m)
    tab["kernels"] = [k for k in tab["kernels"] if not (k["kernel"] == row["kernel"] and k["envs"] == envs and k["dtype"] == dtype)] + [row]
    print(row)
json.dump(tab, open(out_path, "w"), indent=1)
