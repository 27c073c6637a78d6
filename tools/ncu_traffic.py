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
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[hdr.index("Kernel Name")]
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(m)
        tot += float(r[i].replace(",", "")) * scale[units[i]]
    dur = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
    row = {"kernel": name.split("(")[0].replace("void ", ""), "envs": envs, "dtype": dtype, "dram_bytes_per_launch": tot,
           "ncu_duration_" + units[hdr.index("gpu__time_duration.sum")]: dur, "report": os.path.basename(rep)}
    tab["kernels"] = [k for k in tab["kernels"] if not (k["kernel"] == row["kernel"] and k["envs"] == envs and k["dtype"] == dtype)] + [row]
    print(row)
json.dump(tab, open(out_path, "w"), indent=1)
