"""Executed warp-instruction mix by SASS opcode from an ncu report: python tools/ncu_opcodes.py report.ncu-rep [kernel substring]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(starts[:-1], starts[1:]):
    if want in rows[a][1]:
        rows = rows[a:b]; break
hdr = rows[1]
si, ni = hdr.index("Source"), hdr.index("Instructions Executed")
mix = collections.Counter()
for r in rows[2:]:
    if len(r) != len(hdr): continue
    toks = r[si].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    mix[op.split(".")[0]] += int(r[ni] or 0)
tot = sum(mix.values())
print(rows[0][1][:90], "total", tot)
for k, v in mix.most_common(40):
    print(f"  {k:<10} {100 * v / tot:6.2f}%")
