"""The "library kernels to beat" (SURVEY.md section 8d): the same ConvLSTM training step written with stock PyTorch
ops (nn.Conv2d -> cuDNN, eager pointwise kernels, autograd, torch.optim.Adam) on the same B200, BASELINE cfg 2.
Self-contained on purpose: it neither imports the reference (absent on the GPU box) nor `oracle/` (test
infrastructure).  Run on the GPU box:  python tools/torch_eager_baseline.py [--batch 32]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402


class Cell(nn.Module):
    def __init__(self, cin, hc, k):
        super().__init__()
        self.hc = hc
        self.conv = nn.Conv2d(cin + hc, 4 * hc, k, padding=k // 2)

    def forward(self, x, h, c):
        i, f, g, o = torch.split(self.conv(torch.cat([x, h], 1)), self.hc, 1)
        c = c * torch.sigmoid(f) + torch.sigmoid(i) * torch.tanh(g)
        return torch.sigmoid(o) * torch.tanh(c), c


class Net(nn.Module):
    def __init__(self, cin, hc, k):
        super().__init__()
        self.cell, self.head, self.hc = Cell(cin, hc, k), nn.Conv2d(hc, 1, 1), hc

    def forward(self, x):
        B, T, _, H, W = x.shape
        h = torch.zeros(B, self.hc, H, W, device=x.device)
        c = torch.zeros_like(h)
        for t in range(T):
            h, c = self.cell(x[:, t], h, c)
        return self.head(h)


def run(batch, T, mode, steps=5):
    torch.manual_seed(0)
    net = Net(21, 64, 3).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.5, 0.999), fused=True)
    x, y = torch.randn(batch, T, 21, 90, 144, device="cuda"), torch.randn(batch, 90, 144, device="cuda")
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.benchmark = True

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16-autocast"):
            p = net(x).squeeze(1).float()
        loss = F.mse_loss(p, y) + F.l1_loss(p, y)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"torch eager + cuDNN, {mode}: {ms:.1f} ms/step, {batch / ms * 1e3:.0f} samples/s "
          f"(B={batch}, T={T}, 21->64 ch, k3, 90x144; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seq-len", type=int, default=12)
    a = ap.parse_args()
    print(f"torch {torch.__version__}, cuDNN {torch.backends.cudnn.version()}, {torch.cuda.get_device_name(0)}")
    for mode in ("fp32", "tf32", "bf16-autocast"):
        run(a.batch, a.seq_len, mode)
