"""Re-run one fuzz case with per-parameter errors: python tests/tools/fuzz_one.py B T C H W "h1,h2" "k1,k2" precision seed"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nasa_niswan_b200 import ConvLSTM  # noqa: E402
from oracle import convlstm_oracle as O  # noqa: E402

B, T, C, H, W = (int(v) for v in sys.argv[1:6])
hidden = [int(v) for v in sys.argv[6].split(",")]
ks = [int(v) for v in sys.argv[7].split(",")]
precision, seed = sys.argv[8], int(sys.argv[9])
torch.manual_seed(seed)
L = len(hidden)
net = ConvLSTM(C, hidden, ks, L, precision=precision)
params = {k: v.detach().clone() for k, v in net.state_dict().items()}
net = net.cuda()
x = torch.randn(B, T, C, H, W)
leaf = {k: v.clone().requires_grad_(True) for k, v in params.items()}
rp = O.convlstm_forward(x, leaf, L)
y = torch.randn(B, H, W)
dpred = torch.autograd.grad(O.training_loss(rp, y), rp, retain_graph=True)[0]
rp.backward(dpred)
pred = net(x.cuda())
pred.backward(dpred.cuda())
print("pred", O.max_abs_normalised(pred.detach().cpu(), rp.detach()))
for k, p in net.named_parameters():
    g, r = p.grad.cpu(), leaf[k].grad
    print(k, tuple(g.shape), "err %.3e" % O.max_abs_normalised(g, r), "max|ref| %.3e" % r.abs().max().item())
    if g.dim() == 4 and k.startswith("layers."):
        l = int(k.split(".")[1])
        cin = C if l == 0 else hidden[l - 1]
        den = r.abs().max()
        print("   x-part %.3e  h-part %.3e (both / max|ref| of the tensor);  max|ref| x %.2e h %.2e" % (
            (g[:, :cin] - r[:, :cin]).abs().max() / den, (g[:, cin:] - r[:, cin:]).abs().max() / den,
            r[:, :cin].abs().max(), r[:, cin:].abs().max()))
