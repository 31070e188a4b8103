"""Timeline of CTA 0 of one conv launch (run on the GPU box with NINT_DEBUG_FLAGS=8 [+1/+2]).

    NINT_DEBUG_FLAGS=8 python tools/trace_report.py fwd|bwd

Roles: 0 activation producer (warp 0): (wait start, wait end, issued) per chunk; 1 MMA issuer: tempty wait start per
tile, then (a_full wait start, end) per chunk; 2 epilogue loader: (wait start, wait end, issued) per group; 3 storer:
(wait start, wait end, done) per group; 4 math warp 8: (e_full wait start, end, tfull end, done) per group."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the kernels' experiment knobs exist only in the -DNINT_KNOBS=1 build of the library
_KNOBS = os.path.join(ROOT, "nasa_niswan_b200", "libnint_knobs.so")
if "NINT_LIB" not in os.environ:
    if not os.path.exists(_KNOBS):
        import subprocess
        subprocess.run([sys.executable, "-m", "nasa_niswan_b200.build", "--knobs"], cwd=ROOT, check=True)
    os.environ["NINT_LIB"] = _KNOBS
import torch  # noqa: E402
from nasa_niswan_b200 import Plan, _lib  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B, T, C, H, W, hc, k = 32, 3, 21, 90, 144, 64, 3
torch.manual_seed(0)
plan = Plan(B, T, H, W, C, [hc], [k], precision="bf16", training=True)
plan.set_weights(0, torch.randn(4 * hc, C + hc, k, k, device="cuda") * 0.05, torch.zeros(4 * hc, device="cuda"))
plan.set_head(torch.randn(1, hc, 1, 1, device="cuda"), torch.zeros(1, device="cuda"))
x = torch.randn(B, T, C, H, W, device="cuda")
lib = _lib.load()
N = 8 * 1024
buf = (ctypes.c_longlong * N)()
for _ in range(2):
    pred, _s = plan.forward(x)          # last launch traced = cell step t = T-1 (both K segments)
    if which == "fwd":
        _lib.check(lib.nint_debug_read_trace(buf, N, 1), "trace")
if which == "bwd":
    _lib.check(lib.nint_debug_read_trace(buf, N, 1), "trace")   # drop the forward's stamps
    plan.backward(torch.randn_like(pred))  # last conv launch = dgrad step t = 0
    _lib.check(lib.nint_debug_read_trace(buf, N, 1), "trace")
tr = [[v for v in buf[r * 1024:(r + 1) * 1024] if v] for r in range(8)]
t0 = min(v[0] for v in tr if v)
names = {0: "A-producer", 1: "MMA", 2: "epi-loader", 3: "epi-storer", 4: "math w8", 5: "W-producer"}
per = {0: 3, 1: 3, 2: 3, 3: 3, 4: 4}
for r, v in enumerate(tr):
    if not v:
        continue
    rel = [x_ - t0 for x_ in v]
    print(f"== role {r} {names.get(r)}: {len(v)} stamps, span {rel[-1] - rel[0]} cycles")
    if r in per:
        n = per[r]
        rows = [rel[i:i + n] for i in range(0, len(rel) - n + 1, n)]
        for i, row in enumerate(rows[:14] + rows[-4:]):
            d = [row[j + 1] - row[j] for j in range(n - 1)]
            print(f"   #{i if i < 14 else len(rows) - 18 + i:3d} start {row[0]:8d}  deltas {d}")
        if len(rows) > 3:
            period = (rows[-1][0] - rows[2][0]) / max(len(rows) - 3, 1)
            print(f"   steady period {period:.0f} cycles per item; mean deltas "
                  f"{[round(sum(rw[j + 1] - rw[j] for rw in rows[2:]) / len(rows[2:])) for j in range(n - 1)]}")
    else:
        print("   first 60:", rel[:60])
        print("   last 10:", rel[-10:])
