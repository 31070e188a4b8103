"""One or two native training steps of a given geometry, for `ncu --metrics gpu__time_duration.sum` launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \\
        python tools/launch_list.py B T C H W "h1,h2" "k1,k2" [precision]
tools/launch_summary.py out.csv prints the per-launch durations of the last step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nasa_niswan_b200 import ConvLSTM  # noqa: E402
from nasa_niswan_b200.parallel import Trainer  # noqa: E402

B, T, C, H, W = (int(v) for v in sys.argv[1:6])
hidden = [int(v) for v in sys.argv[6].split(",")]
ks = [int(v) for v in sys.argv[7].split(",")]
precision = sys.argv[8] if len(sys.argv) > 8 else "bf16"
torch.manual_seed(0)
net = ConvLSTM(C, hidden, ks, len(hidden), precision=precision).cuda()
tr = Trainer(net, lr=1e-3, betas=(0.5, 0.999))
x = torch.randn(B, T, C, H, W, device="cuda")
y = torch.randn(B, H, W, device="cuda")
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
