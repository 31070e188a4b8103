"""HBM roofline of the memory-bound kernels (north-star item 4): every layout / pointwise kernel
of the path at BASELINE cfg-2 sizes, timed alone with CUDA events (L2 flushed between repetitions), achieved GB/s =
ALGORITHMIC bytes / time against the measured copy bandwidth in MEASURED_PEAKS.json.

    python tools/pointwise_bench.py [--json out.json]         # on the GPU box
    ncu --set full -k regex:'fuse_bank|pack_cl|head_|loss_mse|adam_dev|fuse_inputs' ... python tools/pointwise_bench.py --once
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nasa_niswan_b200 import ConvLSTM, Plan, _lib  # noqa: E402
from nasa_niswan_b200.preprocess import FrameBank, fuse_inputs  # noqa: E402

B, T, C, H, W, HC = 32, 12, 21, 90, 144, 64


def timed(fn, reps, flush, inner=4):
    """median seconds per call; `inner` back-to-back calls share one event pair so the host's launch latency after the
    first event is amortised (every call streams more than the 126 MB L2, so back-to-back calls do not reuse it)"""
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.add_(1.0)                      # 512 MB write: evicts the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json")
    ap.add_argument("--once", action="store_true", help="one repetition per kernel (for ncu)")
    a = ap.parse_args()
    reps = 1 if a.once else 9
    if a.once:
        global timed
        _t = timed
        timed = lambda fn, reps, flush, inner=1: _t(fn, reps, flush, 1)
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    flush = torch.zeros(128 << 20, device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    rows = []

    def report(name, secs, bytes_, note=""):
        gbs = bytes_ / secs / 1e9
        rows.append({"kernel": name, "us": round(secs * 1e6, 1), "algorithmic_MB": round(bytes_ / 1e6, 1),
                     "GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 3), "note": note})
        print(f"{name:34s} {secs * 1e6:9.1f} us  {bytes_ / 1e6:9.1f} MB  {gbs:8.1f} GB/s  {gbs / peak:6.3f} of {peak:.0f}  {note}", flush=True)

    torch.manual_seed(0)
    vp, st = _lib.ptr, _lib.stream_ptr(dev)
    # ---- preprocessing: raw fields -> frame bank (one kernel) vs the round-1 two-pass path
    N, L, Hp, Wp = B * T, 20, 100, 154
    lev, emis = torch.randn(N, L, H, W, device=dev), torch.randn(N, H, W, device=dev).abs()
    mean, std = torch.randn(L + 1, device=dev), torch.rand(L + 1, device=dev) + 0.5
    s = timed(lambda: FrameBank.from_fields(lev, emis, mean, std, (Hp, Wp), "reference_rnn"), reps, flush)
    report("fuse_bank_kernel<bf16,32>", s, N * (C * H * W * 4 + Hp * Wp * 32 * 2), "raw fp32 fields -> bf16 NHWC bank with halo, 384 frames 90x144 -> 100x154")
    s = timed(lambda: FrameBank.from_fields(lev, emis, mean, std, None, "reflect"), reps, flush)
    report("fuse_bank_kernel<bf16,32> no halo", s, N * (C * H * W * 4 + H * W * 32 * 2), "same, 90x144 grid kept")
    s = timed(lambda: fuse_inputs(lev, emis, mean, std, (Hp, Wp), "reference_rnn"), reps, flush)
    report("fuse_inputs_kernel (fp32 NCHW)", s, N * (C * H * W * 4 + C * Hp * Wp * 4), "round-1 kernel: fp32 NCHW out (needs a second packing pass)")
    del lev, emis
    # ---- input packing: x [B,T,C,H,W] -> X channels-last bf16 (fp32 and bf16 sources), frames -> bank
    x = torch.randn(B, T, C, H, W, device=dev)
    x16 = x.to(torch.bfloat16)
    net = ConvLSTM(C, [HC], [3], 1, precision="bf16").to(dev)
    plan = net.plan_for(x, True)
    X = torch.empty(T * B * H * W * 32, dtype=torch.bfloat16, device=dev)
    s = timed(lambda: FrameBank.from_frames(x.view(B * T, C, H, W), "bf16"), reps, flush)
    report("pack_cl_kernel<float,bf16,32>", s, B * T * H * W * (C * 4 + 32 * 2), "fp32 NCHW -> bf16 NHWC (what nint_forward does with x)")
    s = timed(lambda: FrameBank.from_frames(x16.view(B * T, C, H, W), "bf16"), reps, flush)
    report("pack_cl_kernel<bf16,bf16,32>", s, B * T * H * W * (C * 2 + 32 * 2), "bf16 NCHW (host-staged windows) -> bf16 NHWC")
    # ---- head forward / backward on one h slot, fused loss, Adam
    # the head kernels have no C-ABI entry of their own: time them through a T=1 inference plan's forward tail instead
    plan1 = Plan(B, 1, H, W, C, [HC], [3], precision="bf16", training=False, device=dev)
    plan1.set_weights(0, net.layers[0].conv.weight, net.layers[0].conv.bias)
    plan1.set_head(net.conv.weight, net.conv.bias)
    x1 = x[:, :1].contiguous()
    plan1.profile(True)
    for _ in range(3):
        flush.add_(1.0)
        plan1.forward(x1)
    torch.cuda.synchronize()
    prof = plan1.profile_read()
    plan1.profile(False)
    report("pack + head_fwd_kernel (other, T=1)", prof["other"][0] * 1e-3 / 3, B * H * W * (C * 4 + 32 * 2 + HC * 2 + 4) + 0,
           "launches of class 'other' in a T=1 forward: weight repack (tiny), input pack, head")
    y = torch.randn(B, H, W, device=dev)
    pred = torch.randn(B, 1, H, W, device=dev)
    dpred, loss, stats = torch.empty_like(pred), torch.empty(1, device=dev), torch.zeros(8, device=dev)
    s = timed(lambda: lib.nint_loss_mse_l1(vp(pred), vp(y), B, H, W, 0, H, 0, W, vp(dpred), vp(loss), vp(stats), st), reps, flush)
    report("loss_mse_l1_kernel", s, B * H * W * 12, "pred r + y r + dpred w (1.7 MB each): launch-latency bound at this size")
    n = 196161
    p_, g_, m_, v_ = (torch.randn(n, device=dev).abs() for _ in range(4))
    state = torch.tensor([0.0, 1e-3, 0.0, 0.0], device=dev)
    s = timed(lambda: lib.nint_adam_step_dev(vp(p_), vp(g_), vp(m_), vp(v_), n, vp(state), 0.5, 0.999, 1e-8, 1.0, st), reps, flush)
    report("adam_tick + adam_dev_kernel", s, n * 28, "p rw, g r, m rw, v rw over 196 161 parameters (0.78 MB): launch-latency bound")
    if a.json:
        json.dump({"peak_hbm_gbs": peak, "rows": rows}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
