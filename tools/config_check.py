"""Runs the other BASELINE.json configurations once on the GPU box and prints their throughput (sanity + perf):
cfg 4 (T=120 inference rollout, B=64), cfg 5 (3 layers x hidden 128, 5x5, 180x288 grid, B=8), the reference's
shipped model (5 -> 64/32/16, k 5/3/3, 100x154, T=48, B=8; launcher.sh:17-25) and cfg 2 with 5x5 kernels."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nasa_niswan_b200 import ConvLSTM  # noqa: E402
from nasa_niswan_b200.parallel import Trainer  # noqa: E402


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def train_cfg(name, B, T, C, H, W, hidden, ks, crop=None, n=3):
    torch.manual_seed(0)
    net = ConvLSTM(C, hidden, ks, len(hidden), precision="bf16").cuda()
    tr = Trainer(net, lr=1e-3, betas=(0.5, 0.999), crop=crop)
    x = torch.randn(B, T, C, H, W, device="cuda")
    yh, yw = (crop[1] - crop[0], crop[3] - crop[2]) if crop else (H, W)
    y = torch.randn(B, yh, yw, device="cuda")
    l0 = float(tr.step(x, y))
    ms = timed(lambda: tr.step(x, y), n)
    plan = net.plan_for(x, True)
    plan.profile(True)
    tr.step(x, y)
    prof = plan.profile_read()
    plan.profile(False)
    l1 = float(tr.step(x, y))
    print("   per kernel class (ms, launches): " + ", ".join(f"{k} {v[0]:.2f}/{v[1]}" for k, v in prof.items()), flush=True)
    print(f"{name}: {ms:.2f} ms/step, {B / ms * 1e3:.0f} samples/s, loss {l0:.4f} -> {l1:.4f}, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del net, tr, x, y
    torch.cuda.empty_cache()


def main():
    which = sys.argv[1:] or ["cfg4", "cfg5", "shipped", "k5"]
    if "cfg4" in which:
        torch.manual_seed(0)
        net = ConvLSTM(21, [64], [3], 1, precision="bf16").cuda()
        x = torch.randn(64, 120, 21, 90, 144, device="cuda")
        with torch.no_grad():
            ms = timed(lambda: net(x), 2)
        print(f"cfg4 inference rollout T=120 B=64: {ms:.1f} ms/rollout, {64 / ms * 1e3:.0f} samples/s, "
              f"{ms / 120 * 1e3:.0f} us per step", flush=True)
        del net, x
        torch.cuda.empty_cache()
    if "k5" in which:
        train_cfg("cfg2 with 5x5 kernels (B=32, T=12)", 32, 12, 21, 90, 144, [64], [5])
    if "shipped" in which:
        train_cfg("reference's shipped model (launcher.sh)", 8, 48, 5, 100, 154, [64, 32, 16], [5, 3, 3], crop=(5, 95, 5, 149))
    if "cfg5" in which:
        train_cfg("cfg5 3x128 k5 180x288 (B=8, T=12)", 8, 12, 21, 180, 288, [128, 128, 128], [5, 5, 5], n=2)


if __name__ == "__main__":
    main()
