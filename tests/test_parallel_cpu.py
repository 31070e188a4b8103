"""CPU, world_size 2 over gloo: the data-parallel host logic (batch sharding, flat gradient buffer,
mean all-reduce, replicated Adam step) with a tiny stand-in module (the CUDA ConvLSTM cannot run here)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nasa_niswan_b200.parallel import FlatGradients, Trainer, shard_batch


def test_shard_batch_partitions():
    for gb, world in [(256, 8), (256, 2), (10, 4), (3, 4), (32, 1)]:
        spans = [shard_batch(gb, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert [shard_batch(256, r, 8) for r in (0, 7)] == [(0, 32), (224, 256)]


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 1, 3, padding=1)

    def forward(self, x):                      # [B,T,C,H,W] -> [B,1,H,W]
        return self.conv(x[:, -1])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)              # different init per rank: broadcast must fix it
    model = _Tiny()
    tr = Trainer(model, lr=1e-2)
    torch.manual_seed(0)
    X, Y = torch.randn(8, 2, 3, 6, 7), torch.randn(8, 6, 7)
    a, b = shard_batch(8, rank, world)
    for _ in range(3):
        tr.step(X[a:b], Y[a:b])
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({"params": gathered}, out)
    dist.destroy_process_group()


def test_two_rank_training_matches_single_process(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)["params"]
    assert torch.allclose(got[0], got[1], atol=0, rtol=0)          # replicas stay identical
    # single process on the whole batch = same math (mean of shard means, equal shard sizes)
    torch.manual_seed(100)
    model = _Tiny()
    tr = Trainer(model, lr=1e-2)
    torch.manual_seed(0)
    X, Y = torch.randn(8, 2, 3, 6, 7), torch.randn(8, 6, 7)
    for _ in range(3):
        tr.step(X, Y)
    ref = torch.cat([p.detach().flatten() for p in model.parameters()])
    assert torch.allclose(got[0], ref, atol=1e-5)


def test_flat_gradients_are_views():
    m = _Tiny()
    fg = FlatGradients(m.parameters())
    m(torch.randn(2, 1, 3, 4, 4)).sum().backward()
    assert fg.flat.abs().sum() > 0
    assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in m.parameters())
    fg.zero()
    assert all(float(p.grad.abs().sum()) == 0 for p in m.parameters())


def test_step_lr_matches_torch_scheduler():
    """parallel.StepLR follows torch.optim.lr_scheduler.StepLR (train.py:72,120; launcher.sh:27 `10 0.9`) epoch by epoch"""
    import torch
    from nasa_niswan_b200.parallel import StepLR

    class Opt:   # the only thing a scheduler touches
        def __init__(self, lr):
            self.param_groups = [{"lr": lr}]

    w = torch.nn.Parameter(torch.zeros(1))
    ref_opt = torch.optim.Adam([w], lr=1e-3)
    ref = torch.optim.lr_scheduler.StepLR(ref_opt, step_size=10, gamma=0.9)
    mine_opt = Opt(1e-3)
    mine = StepLR(mine_opt, 10, 0.9)
    assert mine.get_last_lr() == ref.get_last_lr()
    for _ in range(35):
        ref_opt.step()
        ref.step()
        mine.step()
        assert abs(mine.get_last_lr()[0] - ref.get_last_lr()[0]) < 1e-15
        assert mine_opt.param_groups[0]["lr"] == mine.get_last_lr()[0]
    # resume: like torch, the optimizer's state dict carries the current lr, the scheduler's the epoch counter
    resumed = StepLR(Opt(mine_opt.param_groups[0]["lr"]), 10, 0.9)
    resumed.load_state_dict(mine.state_dict())
    resumed.step()
    mine.step()
    assert resumed.get_last_lr() == mine.get_last_lr()


def test_step_lr_keeps_an_overridden_learning_rate_like_torch():
    """utils.load_checkpoint writes an lr into param_groups (its `lr` argument or the checkpoint's
    `learning_rate`, utils.py:42-48).  torch's StepLR is chainable -- it multiplies the CURRENT lr -- so the override
    survives; a closed-form `base_lr * gamma ** (epoch // step)` would silently go back to the constructor's lr."""
    import torch
    from nasa_niswan_b200.parallel import StepLR

    class Opt:
        def __init__(self, lr):
            self.param_groups = [{"lr": lr}]

    w = torch.nn.Parameter(torch.zeros(1))
    ref_opt = torch.optim.Adam([w], lr=1e-3)
    ref = torch.optim.lr_scheduler.StepLR(ref_opt, step_size=2, gamma=0.5)
    mine_opt = Opt(1e-3)
    mine = StepLR(mine_opt, 2, 0.5)
    ref_opt.param_groups[0]["lr"] = 5e-4          # what load_checkpoint(..., lr=5e-4) does after the scheduler exists
    mine_opt.param_groups[0]["lr"] = 5e-4
    seen = []
    for _ in range(6):
        ref_opt.step()
        ref.step()
        mine.step()
        assert abs(mine_opt.param_groups[0]["lr"] - ref_opt.param_groups[0]["lr"]) < 1e-15
        seen.append(mine_opt.param_groups[0]["lr"])
    assert seen[0] == 5e-4 and abs(seen[1] - 2.5e-4) < 1e-15 and abs(seen[-1] - 6.25e-5) < 1e-15


class _TwoLayer(torch.nn.Module):     # parameter order of ConvLSTM: layers.0.w, layers.0.b, layers.1.w, layers.1.b, head w, b
    def __init__(self):
        super().__init__()
        self.layers = torch.nn.ModuleList([torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.Conv2d(8, 4, 3, padding=1)])
        self.conv = torch.nn.Conv2d(4, 1, 1)


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _TwoLayer()
    grads = FlatGradients(model.parameters())
    torch.manual_seed(7 + rank)
    grads.flat.copy_(torch.randn_like(grads.flat))
    mine = grads.flat.clone()
    # the order Trainer uses: head first, then the layers top-down, all in flight together
    pending = [grads.all_reduce_bucket_async(i) for i in (2, 1, 0)]
    for w in pending:
        w.wait()
    total = mine.clone()
    dist.all_reduce(total)
    if rank == 0:
        torch.save({"bucketed": grads.flat.clone(), "flat": total,
                    "sizes": [b.numel() for b in grads.buckets],
                    "views": all(p.grad.data_ptr() >= grads.flat.data_ptr() for p in model.parameters())}, out)
    dist.destroy_process_group()


def test_bucketed_async_all_reduce_equals_flat_all_reduce(tmp_path):
    """one bucket per layer (weight + bias) plus the head, reduced asynchronously in the order backward
    finishes them, gives exactly the single flat all-reduce"""
    out = str(tmp_path / "buckets.pt")
    mp.spawn(_bucket_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["sizes"] == [8 * 3 * 9 + 8, 4 * 8 * 9 + 4, 4 + 1]
    assert r["views"]
    assert torch.equal(r["bucketed"], r["flat"])
