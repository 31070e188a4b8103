"""Parity report (GPU box): 12-step rollout on the 90x144 grid vs the fp32 CPU oracle, both
precisions, every output / gradient tensor.  Metric: max|a-b|/max|b| and relative L2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nasa_niswan_b200 import ConvLSTM  # noqa: E402
from oracle import convlstm_oracle as O  # noqa: E402


def report(precision, B, T, C, H, W, hidden, ks, seed=0):
    torch.manual_seed(seed)
    net = ConvLSTM(C, hidden, ks, len(hidden), precision=precision)
    params = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    x, y = torch.randn(B, T, C, H, W), torch.randn(B, H, W)
    # backward parity is measured for the SAME upstream gradient (d loss / d pred at the oracle's pred):
    # L1Loss's sign() is discontinuous, so feeding each side its own pred measures sign flips, not kernels
    leaf = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref_pred = O.convlstm_forward(x, leaf, len(hidden))
    ref_loss = O.training_loss(ref_pred, y)
    dpred = torch.autograd.grad(ref_loss, ref_pred, retain_graph=True)[0]
    ref_pred.backward(dpred)
    ref_grads = {k: v.grad for k, v in leaf.items()}
    ref_pred = ref_pred.detach()
    pred = net(x.cuda())
    loss = O.training_loss(pred, y.cuda())
    pred.backward(dpred.cuda())
    print(f"[{precision}] B{B} T{T} C{C} {H}x{W} hidden{hidden} k{ks}: loss {float(loss):.6f} vs {float(ref_loss):.6f}")
    print(f"    pred                      max-abs-norm {O.max_abs_normalised(pred.detach().cpu(), ref_pred):.3e}  "
          f"rel-L2 {O.rel_l2(pred.detach().cpu(), ref_pred):.3e}")
    for k, p in net.named_parameters():
        print(f"    grad {k:22s} max-abs-norm {O.max_abs_normalised(p.grad.cpu(), ref_grads[k]):.3e}  "
              f"rel-L2 {O.rel_l2(p.grad.cpu(), ref_grads[k]):.3e}")


if __name__ == "__main__":
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    for prec in ("tf32", "bf16"):
        report(prec, 2, 12, 21, 90, 144, [64], [3])
        report(prec, 2, 12, 21, 90, 144, [64], [5])
    report("tf32", 1, 6, 5, 100, 154, [64, 32, 16], [5, 3, 3])
    report("bf16", 1, 6, 5, 100, 154, [64, 32, 16], [5, 3, 3])
