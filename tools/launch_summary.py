"""Per-launch durations of the last training step in an ncu `--metrics gpu__time_duration.sum --csv` log."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
out = []
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    if r[mu] in ("ns", "nsecond"):
        v /= 1e3
    elif r[mu] in ("ms", "msecond"):
        v *= 1e3
    out.append((r[kn].split("(")[0].replace("void nint::", "").replace("nint::", "")[:60], v))
start = [i for i, (n, _) in enumerate(out) if "pack_cl" in n][-1]
for n, v in out[start:]:
    print(f"{v:9.1f} us  {n}")
print(f"{sum(v for _, v in out[start:]):9.1f} us  total of the last step ({len(out) - start} launches, serialised by ncu)")
