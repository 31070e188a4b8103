"""Data-parallel training step for the ConvLSTM (one process per GPU, torch.distributed/NCCL).

Mirrors the reference's inner loop (train.py:96-110): pred = model(X); crop; loss = MSE + L1;
zero_grad; backward; Adam(betas=(0.5, 0.999)).step() -- minus its per-batch host syncs
(train.py:113-114).  The reference has no multi-GPU code; sharding is over samples (SURVEY.md
section 8e): every rank runs the full T loop on its slice of the batch with replicated weights and
the only exchange is one sum all-reduce of a flat fp32 gradient buffer (0.78 MB for the BASELINE
geometry, latency-bound), issued on the current stream right after BPTT.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F


def shard_batch(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the global batch owned by `rank` (remainder to low ranks)."""
    base, rem = divmod(global_batch, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class FlatGradients:
    """All parameter gradients as views into one contiguous fp32 buffer -> a single all-reduce."""

    def __init__(self, params, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=torch.float32, device=ref.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.group)
            if world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.div_(world)


def training_loss(pred: torch.Tensor, y: torch.Tensor, crop: Optional[Tuple[int, int, int, int]] = None):
    """MSELoss(y, pred) + L1Loss(y, pred) on the cropped, squeezed prediction (train.py:74-75,102,105)."""
    if crop is not None:
        y0, y1, x0, x1 = crop
        pred = pred[:, :, y0:y1, x0:x1]
    p = pred.squeeze(1)
    return F.mse_loss(p, y) + F.l1_loss(p, y)


class HostFeeder:
    """Double-buffered host -> device input path: the pinned-memory copy of batch i+1 (train.py:92-93) runs on a
    side stream while batch i trains.  `put` starts the copy and returns a handle, `get` makes the current stream
    wait for it, `release` marks the device buffers reusable once the step that read them has been enqueued."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [None] * depth
        self.i = 0

    def put(self, xh: torch.Tensor, yh: torch.Tensor):
        k = self.i % len(self.slots)
        self.i += 1
        slot = self.slots[k]
        if slot is None or slot["x"].shape != xh.shape or slot["y"].shape != yh.shape:
            slot = {"x": torch.empty(xh.shape, dtype=xh.dtype, device=self.device),
                    "y": torch.empty(yh.shape, dtype=yh.dtype, device=self.device),
                    "ready": torch.cuda.Event(), "free": None}
            self.slots[k] = slot
        with torch.cuda.stream(self.stream):
            if slot["free"] is not None:
                self.stream.wait_event(slot["free"])      # the step that last read this slot has finished
            slot["x"].copy_(xh, non_blocking=True)
            slot["y"].copy_(yh, non_blocking=True)
            slot["ready"].record(self.stream)
        return slot

    @staticmethod
    def get(slot):
        torch.cuda.current_stream().wait_event(slot["ready"])
        return slot["x"], slot["y"]

    @staticmethod
    def release(slot):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        slot["free"] = ev


class Trainer:
    def __init__(self, model, lr: float = 1e-3, betas=(0.5, 0.999), crop=None, process_group=None):
        self.model, self.crop = model, crop
        self.grads = FlatGradients(model.parameters(), process_group)
        fused = next(model.parameters()).is_cuda
        self.optimizer = torch.optim.Adam(self.grads.params, lr=lr, betas=betas, fused=fused)   # train.py:71
        self.broadcast_parameters(process_group)

    def broadcast_parameters(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for p in self.model.parameters():
                dist.broadcast(p.data, src=0, group=group)

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """One optimizer step on this rank's shard; returns the (un-synchronised) local loss tensor."""
        self.grads.zero()                                   # train.py:108
        pred = self.model(x)                                # train.py:96
        loss = training_loss(pred, y, self.crop)            # train.py:102,105
        loss.backward()                                     # train.py:109
        self.grads.all_reduce_mean()
        self.optimizer.step()                               # train.py:110
        return loss.detach()
