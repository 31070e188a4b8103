"""Top sampled SASS instructions of an .ncu-rep source page with their dominant stall reason.
Usage: python tools/ncu_top.py <report.ncu-rep> [N]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
si, ni, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = []
for i, r in enumerate(rows[2:]):
    try:
        n = int(r[ni])
    except (ValueError, IndexError):
        continue
    st = sorted(((int(r[c] or 0), h[c]) for c in cols), reverse=True)[:2]
    data.append((i, n, int(r[ie] or 0), r[si].strip(), st))
tot = sum(d[1] for d in data)
print("total samples", tot, "instructions", len(data))
for d in sorted(data, key=lambda d: -d[1])[:N]:
    print(f"{d[0]:5d} {100 * d[1] / tot:5.1f}% exec={d[2]:9d} {d[4][0][1][6:]:>14s}:{d[4][0][0]:<6d} {d[3][:100]}")
