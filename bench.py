"""Headline benchmark: ConvLSTM training samples/s on the BASELINE geometry (90x144 grid, 20 levels
+ BCB emission = 21 channels, hidden 64, T=12), one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU arm: the oracle port on the host cores

A step = forward + MSE+L1 loss + BPTT + gradient all-reduce + Adam (train.py:96-110) on one batch of
synthetic N(0,1) data of the named shape.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

C_IN, HIDDEN, H, W = 21, 64, 90, 144


def conv_flops(batch, T, k, cin=C_IN, hc=HIDDEN, hw=H * W):
    """Algorithmic (necessary) conv FLOPs per launch class for one training step:
    h_{-1}=0 -> no h segment at t=0; no dx for layer 0; padding channels are not counted."""
    n = 4 * hc
    f_x = 2.0 * hw * n * cin * k * k * batch
    f_h = 2.0 * hw * n * hc * k * k * batch
    fwd = T * f_x + (T - 1) * f_h
    return {"gate_conv_fwd": fwd, "dgrad_gate_bwd": (T - 1) * f_h, "wgrad": fwd}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons, watts = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except (ValueError, IndexError):
                continue
            try:
                watts.append(float(r[2]))
            except (ValueError, IndexError):
                pass
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        watts.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w": watts[len(watts) // 2] if watts else None}


def cpu_oracle_run(batch, T, k, steps, warmup, threads):
    """The CPU arm: oracle port of the reference path (oracle/convlstm_oracle.py), fwd+loss+bwd."""
    from oracle import convlstm_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    params = O.init_params(C_IN, [HIDDEN], [k], seed=0)
    x, y = torch.randn(batch, T, C_IN, H, W), torch.randn(batch, H, W)
    for _ in range(warmup):
        O.forward_backward(x, y, params, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.forward_backward(x, y, params, 1)
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))
    batch = 2
    sps, ms = cpu_oracle_run(batch, args.seq_len, args.ksize, max(1, min(args.steps, 5)), min(args.warmup, 1), threads)
    sample = f"oracle port (torch CPU fp32), batch {batch} x T={args.seq_len} fwd+loss+bwd per step, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": "train samples/sec", "value": round(sps, 4), "unit": "samples/s",
        "n_gpus": args.gpus, "steps": max(1, min(args.steps, 5)), "warmup": min(args.warmup, 1),
        "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ConvLSTM train step, 90x144x(20 levels+BCB)=21ch, hidden 64, k{args.ksize}, "
                               f"T={args.seq_len}; CPU sample batch {batch}"},
        "cpu_baseline": {"value": round(sps, 4), "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(sps, 4), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (weak scaling)")
    ap.add_argument("--seq-len", type=int, default=12)
    ap.add_argument("--ksize", type=int, default=3)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from nasa_niswan_b200 import ConvLSTM, _lib
    from nasa_niswan_b200.parallel import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W_, K_ = max(3, args.warmup), args.steps
    B, T, k = args.batch, args.seq_len, args.ksize
    dev = torch.device("cuda", local)

    torch.manual_seed(0)                                   # identical init on every rank (utils.py:77-88)
    model = ConvLSTM(C_IN, [HIDDEN], [k], 1, precision=args.precision).to(dev)
    trainer = Trainer(model, lr=1e-3, betas=(0.5, 0.999))
    torch.manual_seed(1 + rank)
    x = torch.randn(B, T, C_IN, H, W, device=dev)          # 418 MB at B=32: larger than the 126 MB L2
    y = torch.randn(B, H, W, device=dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    for _ in range(W_):
        trainer.step(x, y)
    plan = model.plan_for(x, True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    plan.profile(True)
    n0 = lib.nint_launch_count(-1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K_):
        loss = trainer.step(x, y)
    e1.record()
    barrier()
    launches = lib.nint_launch_count(-1) - n0
    prof = plan.profile_read()
    plan.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total) / K_
    value = B * world / (ms_step * 1e-3)

    # ---------------- end to end through the public API with host buffers (`e2e`)
    # every step copies its own inputs from pinned host memory (train.py:92-93) and reads its loss back
    # (train.py:113); the copy of batch i+1 overlaps the training of batch i (HostFeeder: what a data loader does)
    from nasa_niswan_b200.parallel import HostFeeder
    xh, yh = x.cpu().pin_memory(), y.cpu().pin_memory()
    del x, y
    e2e_steps = max(3, K_ // 2)
    feeder = HostFeeder(dev)
    for _ in range(2):
        slot = feeder.put(xh, yh)
        float(trainer.step(*feeder.get(slot)))
        feeder.release(slot)
    barrier()
    e0.record()
    slot = feeder.put(xh, yh)
    for i in range(e2e_steps):
        xd, yd = feeder.get(slot)
        loss_dev = trainer.step(xd, yd)
        feeder.release(slot)
        if i + 1 < e2e_steps:
            slot = feeder.put(xh, yh)                       # H2D of the next batch while this one trains
        loss_host = float(loss_dev)                         # train.py:113 loss.item(): D2H + sync
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = B * world / (float(ms_e2e) / e2e_steps * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)   # kernels timed inside a long step
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PF sustained (B200_PROFILING.md)"
        if args.precision == "tf32":
            peak_tf, peak_src = peak_tf / 2, peak_src + " / 2 (tf32 nominal half rate)"
        flops = conv_flops(B, T, k)
        kernels = {}
        for name in ("gate_conv_fwd", "dgrad_gate_bwd", "wgrad"):
            ms, n = prof[name]
            if n:
                kernels[name] = {"launches_per_step": n / K_, "ms_per_step": round(ms / K_, 4),
                                 "avg_launch_us": round(ms / n * 1e3, 2),
                                 "tflops": round(flops[name] * K_ / (ms * 1e-3) / 1e12, 1),
                                 "frac": round(flops[name] * K_ / (ms * 1e-3) / 1e12 / peak_tf, 4)}
        dominant = max(kernels, key=lambda n_: kernels[n_]["ms_per_step"])
        conv_ms = sum(v["ms_per_step"] for v in kernels.values())
        # the fused dgrad + gate-backward kernel is HBM-bound.  Algorithmic bytes per pixel of one launch
        # (bf16 gates / fp32 state, hidden 64): gates_t r 512 + dgates_t w 512 + dc w 256 in every launch; the dgrad
        # operand dgates_{t+1} r 512, c_{t-1} r 256 and dc r 256 in T-1 of the T launches (nothing flows into the last
        # step, nothing precedes the first); c_t is recomputed, not read.  Average over the T launches of a step.
        esz = 2 if args.precision == "bf16" else 4
        gates_b, state_b = 4 * HIDDEN * esz, HIDDEN * 4
        bwd_bytes = int(B * H * W * (2 * gates_b + state_b + (gates_b + 2 * state_b) * (T - 1) / T))
        hbm_peak = peaks.get("hbm_gbs", 6550.0)
        if "dgrad_gate_bwd" in kernels:
            kb = kernels["dgrad_gate_bwd"]
            kb["hbm_gbs"] = round(bwd_bytes / (kb["avg_launch_us"] * 1e-6) / 1e9, 1)
            kb["hbm_frac"] = round(kb["hbm_gbs"] / hbm_peak, 4)
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dominant)
        except (OSError, ValueError):
            pass
        if dominant == "dgrad_gate_bwd":
            roof = {"bound": "hbm", "kernel": dominant, "achieved": kernels[dominant]["hbm_gbs"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": kernels[dominant]["hbm_frac"], "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6550 GB/s (B200_PROFILING.md)",
                    "bytes_per_launch": bwd_bytes}
        else:
            roof = {"bound": "tensor", "kernel": dominant, "achieved": kernels[dominant]["tflops"], "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": kernels[dominant]["frac"], "traffic": traffic, "peak_source": peak_src,
                    "flops_per_launch": flops[dominant] / (kernels[dominant]["launches_per_step"])}
        total_tf = sum(flops.values()) / (conv_ms * 1e-3) / 1e12
        out = {
            "metric": "train samples/sec", "value": round(value, 2), "unit": "samples/s", "n_gpus": world,
            "steps": K_, "warmup": W_, "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"ConvLSTM train step (fwd+MSE/L1+BPTT+allreduce+Adam), 90x144 grid, 20 levels + BCB "
                                   f"= 21 ch, hidden 64, k{k}, T={T}, batch {B}/GPU (global {B * world})",
                       "parallelism": f"dp{world}", "l2_policy": "inputs (418 MB/step at B=32) larger than the 126 MB L2",
                       "loss_at_end": round(float(loss), 5)},
            "e2e": {"value": round(e2e_value, 2), "unit": "samples/s", "h2d_bytes_per_step": xh.numel() * 4 + yh.numel() * 4,
                    "d2h_bytes_per_step": 4, "steps": e2e_steps,
                    "how": "Trainer.step on host batches: pinned fp32 x,y copied every step (double-buffered on a copy "
                           "stream), loss read back every step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernels": kernels,
            "gate_conv_fwd_bwd": {"tflops": round(total_tf, 1), "frac": round(total_tf / peak_tf, 4),
                                  "conv_ms_per_step": round(conv_ms, 3), "other_ms_per_step": round(ms_step - conv_ms, 3)},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = len(os.sched_getaffinity(0))
            sps, ms = cpu_oracle_run(2, T, k, 3, 1, threads)
            out["cpu_baseline"] = {"value": round(sps, 4), "unit": "samples/s", "cores": threads, "kind": "port",
                                   "sample": f"oracle port (torch CPU fp32) fwd+loss+bwd, batch 2 x T={T}, 3 timed steps "
                                             f"of {ms:.0f} ms, {threads} threads"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
