"""A/B timer for the cfg-2 training step without per-launch profiling events (which serialise launches and hide what
programmatic dependent launch or a CUDA graph buy): ms per step over --steps steps, CUDA events around the whole loop.

    NINT_PDL=1 python tools/step_time.py --steps 150 [--bank] [--graph] [--batch 32] [--seq-len 12]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nasa_niswan_b200 import ConvLSTM  # noqa: E402
from nasa_niswan_b200.parallel import Trainer  # noqa: E402
from nasa_niswan_b200.preprocess import FrameBank  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seq-len", type=int, default=12)
    ap.add_argument("--ksize", type=int, default=3)
    ap.add_argument("--shipped", action="store_true", help="the reference's recipe: 5 -> 64/32/16, k 5/3/3, 100x154, T=48, B=8, crop")
    ap.add_argument("--bank", action="store_true")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--profile", action="store_true", help="afterwards: 5 more steps with per-launch events, time per kernel class")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, T, C, H, W = a.batch, a.seq_len, 21, 90, 144
    torch.manual_seed(0)
    crop = None
    if a.shipped:
        B, T, C, H, W, crop = 8, 48, 5, 100, 154, (5, 95, 5, 149)
        net = ConvLSTM(C, [64, 32, 16], [5, 3, 3], 3, precision="bf16").to(dev)
    else:
        net = ConvLSTM(C, [64], [a.ksize], 1, precision="bf16").to(dev)
    tr = Trainer(net, lr=1e-3, betas=(0.5, 0.999), crop=crop)
    if a.bank:
        bank = FrameBank.from_frames(torch.randn(512, C, H, W, device=dev), "bf16", targets=torch.randn(512, H, W, device=dev))
        idx = torch.randint(0, 512 - T + 1, (B,), dtype=torch.int32).to(dev)
        step = lambda: tr.step_windows(bank, idx, T)
        if a.graph:
            tr.capture(bank=bank, win_start=idx, seq_len=T)
            step = tr.replay
    else:
        x = torch.randn(B, T, C, H, W, device=dev)
        y = torch.randn(B, crop[1] - crop[0], crop[3] - crop[2], device=dev) if crop else torch.randn(B, H, W, device=dev)
        step = lambda: tr.step(x, y)
        if a.graph:
            tr.capture(x, y)
            step = tr.replay
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"pdl={os.environ.get('NINT_PDL', '0')} bank={int(a.bank)} graph={int(a.graph)} shipped={int(a.shipped)} B={B} T={T} k{a.ksize}: "
          f"{ms:.3f} ms/step, {B / ms * 1e3:.0f} samples/s over {a.steps} steps, loss {float(loss):.4f}", flush=True)
    if a.profile and not a.bank:
        plan = net.plan_for(x, True)
        plan.profile(True)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        prof = plan.profile_read()
        plan.profile(False)
        print("   per step: " + ", ".join(f"{k} {v[0] / 5:.3f} ms ({v[1] // 5} steps/launches)" for k, v in prof.items()), flush=True)


if __name__ == "__main__":
    main()
