import os, sys
sys.path.insert(0, '/root/repo')
import torch
from nasa_niswan_b200 import ConvLSTM
from nasa_niswan_b200.parallel import Trainer
torch.manual_seed(0)
net = ConvLSTM(5, [64, 32, 16], [5, 3, 3], 3, precision="bf16").cuda()
tr = Trainer(net, lr=1e-3, betas=(0.5, 0.999), crop=(5, 95, 5, 149))
x = torch.randn(8, 4, 5, 100, 154, device="cuda")
y = torch.randn(8, 90, 144, device="cuda")
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
