"""Per-kernel roofline lines of a multi-kernel .ncu-rep (run where ncu is installed; no GPU needed): duration, DRAM
bytes read / written, achieved DRAM GB/s against the measured copy bandwidth, DRAM / L2 / SM busy percentages.
Usage: python tools/ncu_multi.py <report.ncu-rep> [out.txt]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    peak = 6553.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        pass

    def get(r, name, want_unit=None):
        i = col.get(name)
        if i is None:
            return float("nan")
        v, u = num(r[i]), units[i]
        scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
                 "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        return v * scale
    print(f"# {os.path.basename(rep)}: ncu --set full --clock-control none; GB/s = (dram read + write) / duration, "
          f"fraction of the measured copy bandwidth {peak:.0f} GB/s", file=out)
    print(f"{'kernel':58s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>8s} {'frac':>6s} {'dram%':>6s} {'l2%':>6s} {'sm%':>6s} {'regs':>5s}", file=out)
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]].replace("void nint::", "").replace("nint::", "").replace("__nv_bfloat16", "bf16")[:58]
        us = get(r, "gpu__time_duration.sum")
        rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else float("nan")
        print(f"{name:58s} {us:8.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {gbs:8.1f} {gbs / peak:6.3f} "
              f"{get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{get(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{get(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{get(r, 'launch__registers_per_thread'):5.0f}", file=out)


if __name__ == "__main__":
    main()
