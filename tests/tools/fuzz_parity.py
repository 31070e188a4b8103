"""Randomised geometry sweep on the GPU box: forward + BPTT of random small ConvLSTM configurations against the CPU
oracle (the same check as tests/test_gpu_parity.py::test_plans_against_oracle).  Prints one line per case and a
summary; exit code 1 if any case raises or is off by more than 5x the parity bar.

This is a bug finder, not the parity gate: the bars (2e-2 bf16, 1e-3 tf32) are defined for the 12-step rollout on
the 90x144 grid, where gradients are long, coherent sums.  Tiny random cases have ill-conditioned gradients (a bias
gradient is a plain sum of a few hundred signed terms that nearly cancel, so the 2^-9 rounding of the stored gates
shows up amplified: 2e-2..8e-2 in bf16); such cases are listed as "marginal" and reproduce bit-for-bit the same
error in the CTA-pair and the single-CTA kernels (tests/tools/fuzz_one.py), while an indexing bug gives errors of O(1).

    python tests/tools/fuzz_parity.py [n_cases] [seed]
"""
import os
import random
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nasa_niswan_b200 import ConvLSTM  # noqa: E402
from oracle import convlstm_oracle as O  # noqa: E402

BAR = {"bf16": 2e-2, "tf32": 1e-3}      # north-star parity bars
FAIL_FACTOR = 5.0
HIDDEN = [10, 16, 24, 32, 48, 64, 70, 128, 192, 256]      # incl. sizes that run padded (model.py:207 takes any int)


def random_case(rng):
    L = rng.choice([1, 1, 2, 3])
    hidden = [rng.choice(HIDDEN) for l in range(L)]
    ks = [rng.choice([1, 3, 3, 5, 7]) for _ in range(L)]
    C = rng.choice([1, 3, 5, 16, 21, 32, 33, 40, 64])
    seq = rng.random() < 0.25
    return dict(B=rng.randint(1, 3), T=rng.randint(1, 4), C=C, H=rng.randint(2, 40), W=rng.randint(2, 40),
                hidden=hidden, ks=ks, precision=rng.choice(["bf16", "tf32"]), seq=seq,
                # round 2: windows cut out of an HBM-resident frame bank instead of an explicit tensor; gradient w.r.t. x
                bank=(not seq) and rng.random() < 0.3, dx=(not seq) and rng.random() < 0.3)


def run_case(c, seed):
    torch.manual_seed(seed)
    L = len(c["hidden"])
    net = ConvLSTM(c["C"], c["hidden"], c["ks"], L, precision=c["precision"], return_sequence=c["seq"])
    params = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    x = torch.randn(c["B"], c["T"], c["C"], c["H"], c["W"])
    starts = None
    if c.get("bank"):
        record = torch.randn(c["T"] + 3, c["C"], c["H"], c["W"])
        starts = torch.randint(0, 4, (c["B"],))
        x = torch.stack([record[s:s + c["T"]] for s in starts.tolist()])
    leaf = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    if c["seq"]:
        rp, rs = O.convlstm_forward(x, leaf, L, return_sequence=True)
        wts = torch.linspace(0.5, 1.5, c["T"]).view(1, -1, 1, 1)
        ((rs * wts).sum() / rs.numel() + rp.mean()).backward()
        pred, seq = net(x.cuda())
        ((seq * wts.cuda()).sum() / seq.numel() + pred.mean()).backward()
        err = max(O.max_abs_normalised(pred.detach().cpu(), rp.detach()), O.max_abs_normalised(seq.detach().cpu(), rs.detach()))
    else:
        xr = x.clone().requires_grad_(bool(c.get("dx")))
        rp = O.convlstm_forward(xr, leaf, L)
        y = torch.randn(c["B"], c["H"], c["W"])
        dpred = torch.autograd.grad(O.training_loss(rp, y), rp, retain_graph=True)[0]
        rp.backward(dpred)
        if starts is not None:
            from nasa_niswan_b200.preprocess import FrameBank
            pred = net.forward_windows(FrameBank.from_frames(record.cuda(), c["precision"]), starts, c["T"])
            xd = None
        else:
            xd = x.clone().cuda().requires_grad_(bool(c.get("dx")))
            pred = net(xd)
        pred.backward(dpred.cuda())
        err = O.max_abs_normalised(pred.detach().cpu(), rp.detach())
    gerr = max(O.max_abs_normalised(p.grad.cpu(), leaf[k].grad) for k, p in net.named_parameters())
    if not c["seq"] and c.get("dx") and xd is not None and xr.grad.abs().max() > 0:
        gerr = max(gerr, O.max_abs_normalised(xd.grad.cpu(), xr.grad))
    return err, gerr


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = random.Random(seed)
    bad = refused = marginal = 0
    t0 = time.time()
    for i in range(n):
        c = random_case(rng)
        tag = (f"B{c['B']} T{c['T']} C{c['C']} {c['H']}x{c['W']} h{c['hidden']} k{c['ks']} {c['precision']}"
               f"{' seq' if c['seq'] else ''}{' bank' if c.get('bank') else ''}{' dx' if c.get('dx') else ''}")
        try:
            err, gerr = run_case(c, seed * 1000 + i)
            bar = BAR[c["precision"]]
            worst = max(err, gerr)
            verdict = "ok  " if worst < bar else ("marg" if worst < FAIL_FACTOR * bar else "FAIL")
            bad += verdict == "FAIL"
            marginal += verdict == "marg"
            print(f"[{i:3d}] {verdict} pred {err:.2e} grad {gerr:.2e}  {tag}", flush=True)
        except RuntimeError as e:
            msg = str(e)
            # geometry limits are refused when the plan is made: that is the specified behaviour
            if "nint_plan_create" in msg and ("not supported" in msg or "do not fit" in msg or "unsupported" in msg):
                refused += 1
                print(f"[{i:3d}] refused ({msg.split(':', 1)[1].strip()[:90]})  {tag}", flush=True)
            else:
                bad += 1
                print(f"[{i:3d}] ERROR {msg[:300]}  {tag}", flush=True)
                traceback.print_exc()
        torch.cuda.empty_cache()
    print(f"{n} cases: {n - bad - marginal - refused} within the parity bar, {marginal} marginal (< {FAIL_FACTOR:g}x the bar), "
          f"{bad} bad, {refused} refused by plan validation, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
