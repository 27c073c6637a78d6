"""Top SASS instructions by a stall reason from an ncu report's source page.
usage: python tools/ncu_stalls.py report.ncu-rep stall_long_sb [kernel substring] [top_n]"""
import csv, subprocess, sys
rep, col = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else ""
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(starts[:-1], starts[1:]):
    if want in rows[a][1]:
        rows = rows[a:b]; break
hdr = rows[1]
ci, si, ni = hdr.index(col), hdr.index("Source"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ci] or 0) for r in body)
print(col, "total samples", tot)
order = sorted(range(len(body)), key=lambda i: -int(body[i][ci] or 0))[:top]
for i in order:
    ctx = " | ".join(body[j][si].strip()[:48] for j in range(max(0, i - 2), i))
    print(f"{int(body[i][ci] or 0):6d}  @{body[i][0]}  {body[i][si].strip()[:70]}   <= {ctx}")
