"""Data-parallel training step for the ConvLSTM (one process per GPU, torch.distributed/NCCL).

Mirrors the reference's inner loop (train.py:96-110): pred = model(X); crop; loss = MSE + L1;
zero_grad; backward; Adam(betas=(0.5, 0.999)).step() -- minus its per-batch host syncs
(train.py:113-114).  The reference has no multi-GPU code; sharding is over samples (SURVEY.md
section 8e): every rank runs the full T loop on its slice of the batch with replicated weights and
the only exchange is one sum all-reduce of a flat fp32 gradient buffer (0.78 MB for the BASELINE
geometry, latency-bound), issued on the current stream right after BPTT.
"""
import os
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

        # buckets for the overlapped all-reduce: one per ConvLSTM layer (weight + bias are adjacent in
        # model.parameters() order) and one for the 1x1 head
        self.buckets = []
        off = 0
        for i in range(0, len(self.params), 2):
            n = sum(p.numel() for p in self.params[i:i + 2])
            self.buckets.append(self.flat[off:off + n])
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce_bucket_async(self, i: int):
        """Start the sum all-reduce of bucket i (layer i; the last bucket is the head).  NCCL runs it on its own
        stream after the work queued so far on the current one, so kernels launched next overlap with it."""
        return dist.all_reduce(self.buckets[i], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

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


class SymmetricGradients:
    """The flat gradient buffer as a SYMMETRIC allocation (torch.distributed._symmetric_memory: same layout on every
    rank, mapped into every peer's address space over NVLink / NVSwitch): two gradient slots used alternately by step
    parity plus a flag block.  `nint_dp_allreduce_adam` then does the cross-rank barrier, the sum of all ranks' slots in
    rank order and the Adam update in ONE kernel that reads the peers' gradients directly -- no NCCL call, no second pass
    over the gradients, bit-identical sums on every rank."""

    def __init__(self, n: int, device, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        self.n = n
        self.n_pad = (n + 63) // 64 * 64                 # slot size in floats: slots stay 256-byte aligned
        self.buf = symm.empty(2 * self.n_pad + 64, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.rank, self.world = self.hdl.rank, self.hdl.world_size
        if self.world > 16:
            raise RuntimeError("the fused NVLink step tail handles up to 16 ranks")
        self.peer_ptrs = (ctypes.c_void_p * self.world)(*[int(ptr) for ptr in self.hdl.buffer_ptrs])
        self.flags_offset_bytes = 2 * self.n_pad * 4
        self.seq = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)                              # every rank's zeroed flags are in place before the first signal

    def slot(self, parity: int) -> torch.Tensor:
        return self.buf[parity * self.n_pad: parity * self.n_pad + self.n]


def bind_to_gpu_numa_node(device) -> Optional[int]:
    """Pins this process (and therefore the first-touch placement of the pinned staging buffers it allocates next) to
    the CPUs of the NUMA node the GPU hangs off.  With 8 ranks per box all staging from node 0, host-to-device copies
    share one socket's memory controllers and inter-socket links; bound, every rank streams from local DRAM through its
    own PCIe root.  Returns the node, or None when the topology cannot be read (nothing is changed then)."""
    try:
        dev = torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        props = torch.cuda.get_device_properties(idx)
        bus, dom, devid = getattr(props, "pci_bus_id", None), getattr(props, "pci_domain_id", 0), getattr(props, "pci_device_id", 0)
        if bus is None:
            return None
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


class HostFeeder:
    """Double-buffered host -> device input path: the pinned-memory copy of batch i+1 (train.py:92-93) runs on a
    side stream while batch i trains.  `put` starts the copies of any number of host tensors and returns a handle, `get`
    makes the current stream wait for them and returns the device tensors, `release` marks the device buffers reusable
    once the step that read them has been enqueued.  Host tensors keep their dtype: windows staged as bf16 cross PCIe
    at half the bytes of fp32 and give identical results (the model rounds fp32 inputs to bf16 anyway)."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [None] * depth
        self.i = 0

    def put(self, *host_tensors):
        k = self.i % len(self.slots)
        self.i += 1
        slot = self.slots[k]
        if slot is None or [(t.shape, t.dtype) for t in slot["dev"]] != [(t.shape, t.dtype) for t in host_tensors]:
            slot = {"dev": [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors],
                    "ready": torch.cuda.Event(), "free": None}
            self.slots[k] = slot
        with torch.cuda.stream(self.stream):
            if slot["free"] is not None:
                self.stream.wait_event(slot["free"])      # the step that last read this slot has finished
            for d, h in zip(slot["dev"], host_tensors):
                d.copy_(h, non_blocking=True)
            slot["ready"].record(self.stream)
        return slot

    def get(self, slot):
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        return tuple(slot["dev"])

    def release(self, slot):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        slot["free"] = ev


class NativeAdam:
    """torch.optim.Adam(lr, betas) (train.py:71) as ONE kernel over flat buffers (nint_adam_step_dev).  The parameters
    are re-pointed at views of one flat fp32 buffer (names, shapes and values unchanged: `state_dict` is untouched);
    `state_dict` / `load_state_dict` speak torch.optim.Adam's format so utils.py:23-50 checkpoints interoperate.
    The step count and the learning rate live in device memory (`self.state` = {step, lr, ...}): no launch argument
    changes from step to step, so a CUDA graph of the whole training step can be replayed."""

    def __init__(self, params, flat_grads: torch.Tensor, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        from . import _lib
        self.lib = _lib.load()
        self.params = list(params)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.device = dev
        self.flat_params = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                view = self.flat_params[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                off += p.numel()
        self.flat_grads = flat_grads
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self.param_groups = [{"lr": self.lr, "betas": self.betas, "eps": self.eps}]   # what LR schedulers touch
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)                  # {step, lr, bc1, sqrt(bc2)}
        self._lr_on_device = None
        self._sync_lr()

    def _sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_device:        # schedulers change it once per epoch: a 4-byte upload then, nothing per step
            self.state[1:2].copy_(torch.tensor([lr], dtype=torch.float32), non_blocking=False)
            self._lr_on_device = lr

    def step(self, grad_scale: float = 1.0):
        from . import _lib
        self.step_count += 1
        self._sync_lr()
        vp = _lib.ptr
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_adam_step_dev(vp(self.flat_params), vp(self.flat_grads), vp(self.exp_avg),
                                                   vp(self.exp_avg_sq), self.flat_params.numel(), vp(self.state),
                                                   self.betas[0], self.betas[1], self.eps, float(grad_scale),
                                                   _lib.stream_ptr(self.device)), "nint_adam_step_dev")
        for p in self.params:      # modified in place behind torch's back: keep the version counters honest
            torch.autograd.graph.increment_version(p)

    def state_dict(self):
        state, off = {}, 0
        step = int(round(float(self.state[0])))     # the device counter is the truth (graph replays advance it)
        for i, p in enumerate(self.params):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(step)),
                        "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
            off += n
        group = {"lr": float(self.param_groups[0]["lr"]), "betas": self.betas, "eps": self.eps, "weight_decay": 0,
                 "amsgrad": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        off = 0
        for i, p in enumerate(self.params):
            n = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                self.step_count = int(float(st["step"]))
            off += n
        self.state[0:1].copy_(torch.tensor([float(self.step_count)], dtype=torch.float32))
        g = sd["param_groups"][0]
        self.param_groups[0]["lr"] = float(g["lr"])
        self.betas, self.eps = (float(g["betas"][0]), float(g["betas"][1])), float(g["eps"])
        self._sync_lr()


class StepLR:
    """torch.optim.lr_scheduler.StepLR(optimizer, step_size, gamma) (train.py:72) for any optimizer that exposes
    `param_groups` -- torch's own class insists on a torch.optim.Optimizer, which NativeAdam is not.  Stepped once per
    epoch (train.py:120).  Same chainable rule as torch: every `step_size`-th epoch the CURRENT learning rate of each
    group is multiplied by gamma, so a rate written into `param_groups` by utils.load_checkpoint (its `lr` argument or
    the checkpoint's `learning_rate`, utils.py:42-48) is kept, exactly as with torch's scheduler."""

    def __init__(self, optimizer, step_size, gamma=0.1):
        self.optimizer, self.step_size, self.gamma = optimizer, int(step_size), float(gamma)
        if self.step_size < 1:
            raise ValueError("step_size must be a positive integer")
        self.base_lrs = [float(g.setdefault("initial_lr", g["lr"])) for g in optimizer.param_groups]
        self.last_epoch = 0
        self._last_lr = [float(g["lr"]) for g in optimizer.param_groups]

    def step(self):
        self.last_epoch += 1
        if self.last_epoch % self.step_size == 0:
            for g in self.optimizer.param_groups:
                g["lr"] = float(g["lr"]) * self.gamma
        self._last_lr = [float(g["lr"]) for g in self.optimizer.param_groups]

    def get_last_lr(self):
        return list(self._last_lr)

    def state_dict(self):
        return {"step_size": self.step_size, "gamma": self.gamma, "base_lrs": list(self.base_lrs),
                "last_epoch": self.last_epoch, "_last_lr": list(self._last_lr)}

    def load_state_dict(self, sd):
        self.step_size, self.gamma = int(sd["step_size"]), float(sd["gamma"])
        self.base_lrs, self.last_epoch = [float(v) for v in sd["base_lrs"]], int(sd["last_epoch"])
        self._last_lr = [float(v) for v in sd["_last_lr"]]


class Trainer:
    """`native=True` (default on CUDA): the whole step stays in libnint kernels -- forward, fused MSE+L1 loss with its
    gradient (nint_loss_mse_l1), BPTT writing straight into the flat all-reduce buffer, one Adam kernel -- with no
    autograd graph and ~25 fewer small launches per step.  `native=False` keeps torch's loss / autograd / fused Adam
    around the same ConvLSTM kernels (identical numerics up to summation order).

    Three ways to feed a step (all the same arithmetic):
      step(x, y)                      x [B,T,C,H,W] fp32 or bf16 on the device (train.py:92-96)
      step_windows(bank, win_start)   window start indices into an HBM-resident preprocess.FrameBank (dataset.py:551-637)
      capture(...) / replay()         the native step as ONE CUDA graph over static input buffers"""

    def __init__(self, model, lr: float = 1e-3, betas=(0.5, 0.999), crop=None, process_group=None, native=None,
                 scheduler_config=None):
        self.model, self.crop, self.group = model, crop, process_group
        # one flat all-reduce after backward (default) or bucketed all-reduces overlapped with the wgrad kernels
        # (NINT_DP_OVERLAP=1).  Measured at 2 x B200, cfg 2: overlapping is SLOWER (6.71-6.82 vs 6.64-6.67 ms/step, wgrad
        # 2.02 vs 1.83 ms) -- the wgrad kernel is persistent with one CTA per SM, so NCCL's CTAs take SMs its static
        # tile partition counts on, and a 0.78 MB all-reduce costs ~30 us when it runs alone.
        self.overlap = os.environ.get("NINT_DP_OVERLAP", "0") == "1"
        self.grads = FlatGradients(model.parameters(), process_group)
        on_cuda = next(model.parameters()).is_cuda
        self.device = next(model.parameters()).device
        self.native = on_cuda if native is None else bool(native)
        if self.native:
            self.optimizer = NativeAdam(self.grads.params, self.grads.flat, lr=lr, betas=betas)
            self._stats = torch.zeros(8, dtype=torch.float32, device=self.grads.flat.device)
        else:
            self.optimizer = torch.optim.Adam(self.grads.params, lr=lr, betas=betas, fused=on_cuda)   # train.py:71
        # train.py:72 / launcher.sh:27: `--scheduler-config STEP GAMMA`, stepped once per epoch by `end_epoch`
        self.scheduler = None if scheduler_config is None else StepLR(self.optimizer, int(scheduler_config[0]),
                                                                      float(scheduler_config[1]))
        self._graph = None
        self.broadcast_parameters(process_group)
        # multi-GPU step tail: one fused kernel over NVLink peer memory (barrier + all-reduce + Adam) instead of
        # ncclAllReduce + the Adam kernel; NINT_DP_FUSED=0, or a box without peer mappings, keeps NCCL
        self.sym = None
        if self.native and self._world() > 1 and not self.overlap and os.environ.get("NINT_DP_FUSED", "1") == "1":
            try:
                self.sym = SymmetricGradients(self.grads.flat.numel(), self.device, process_group)
                self._sym_views = [self._views_into(self.sym.slot(0)), self._views_into(self.sym.slot(1))]
            except Exception as exc:     # noqa: BLE001  (no symmetric-memory support here: say so once, use NCCL)
                import warnings
                warnings.warn(f"fused NVLink step tail unavailable ({type(exc).__name__}: {exc}); using NCCL all-reduce")
                self.sym = None

    def end_epoch(self):
        """train.py:120: `scheduler.step()` after the last batch of an epoch; returns the learning rate(s) now in force."""
        if self.scheduler is not None:
            self.scheduler.step()
            if self.native:
                self.optimizer._sync_lr()      # outside any graph replay: the device copy of lr follows the schedule
            return self.scheduler.get_last_lr()
        return [float(g["lr"]) for g in self.optimizer.param_groups]

    def broadcast_parameters(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for p in self.model.parameters():
                dist.broadcast(p.data, src=0, group=group)

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _grad_views(self):
        ps = self.grads.params            # layers.{l}.conv.weight, .bias, ..., conv.weight, conv.bias (model.parameters() order)
        L = (len(ps) - 2) // 2
        return [ps[2 * l].grad for l in range(L)], [ps[2 * l + 1].grad for l in range(L)], ps[-2].grad, ps[-1].grad

    def _views_into(self, flat: torch.Tensor):
        """the same (weights, biases, head weight, head bias) views laid over another flat buffer"""
        views, off = [], 0
        for p in self.grads.params:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        L = (len(views) - 2) // 2
        return [views[2 * l] for l in range(L)], [views[2 * l + 1] for l in range(L)], views[-2], views[-1]

    def _loss_and_update(self, plan, pred, y, y_index, y_offset):
        """train.py:102-110 after the forward: fused loss + its gradient, BPTT into the flat buffer, all-reduce, Adam."""
        from . import _lib
        B, _, H, W = pred.shape
        y0, y1, x0, x1 = self.crop if self.crop is not None else (0, H, 0, W)
        want = (y1 - y0, x1 - x0)
        if y_index is None and tuple(y.shape) != (B,) + want:
            raise ValueError(f"y has shape {tuple(y.shape)}, expected {(B,) + want}")
        if y_index is not None and tuple(y.shape[1:]) != want:
            raise ValueError(f"target bank has frames of shape {tuple(y.shape[1:])}, expected {want}")
        if y.dtype != torch.float32 or y.device != pred.device:
            raise TypeError("targets must be float32 on the model's device")
        dpred = torch.empty_like(pred)
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        vp, lib = _lib.ptr, _lib.load()
        with _lib.on_device(pred.device):
            st = _lib.stream_ptr(pred.device)
            if y_index is None:
                _lib.check(lib.nint_loss_mse_l1(vp(pred), vp(y.contiguous()), B, H, W, y0, y1, x0, x1, vp(dpred), vp(loss),
                                                vp(self._stats), st), "nint_loss_mse_l1")   # train.py:102,105
            else:
                _lib.check(lib.nint_loss_mse_l1_bank(vp(pred), vp(y), y.shape[0], vp(y_index), int(y_offset), B, H, W, y0, y1, x0, x1,
                                                     vp(dpred), vp(loss), vp(self._stats), st), "nint_loss_mse_l1_bank")
        world = self._world()
        if self.sym is not None:
            # fused tail: BPTT writes this step's slot of the symmetric buffer; one kernel then signals the peers, waits
            # for theirs, sums all ranks' slots over NVLink in rank order and applies Adam
            sym, opt = self.sym, self.optimizer
            sym.seq += 1
            parity = sym.seq & 1
            plan.backward(dpred, out=self._sym_views[parity])
            opt.step_count += 1
            opt._sync_lr()
            with _lib.on_device(pred.device):
                _lib.check(lib.nint_dp_allreduce_adam(sym.peer_ptrs, parity * sym.n_pad * 4, sym.flags_offset_bytes, sym.rank,
                                                      sym.world, sym.seq, vp(opt.flat_params), vp(opt.exp_avg),
                                                      vp(opt.exp_avg_sq), opt.flat_params.numel(), vp(opt.state),
                                                      opt.betas[0], opt.betas[1], opt.eps, 1.0 / world,
                                                      _lib.stream_ptr(pred.device)), "nint_dp_allreduce_adam")
            flat_views = self._sym_views[parity]
            for q, g in zip(opt.params, [t for pair in zip(flat_views[0], flat_views[1]) for t in pair] + [flat_views[2], flat_views[3]]):
                q.grad = g                 # this rank's (un-reduced) gradients of the step; the sum lives only in registers
                torch.autograd.graph.increment_version(q)
            return loss[0]
        if world > 1 and not self.overlap:
            plan.backward(dpred, out=self._grad_views())
            dist.all_reduce(self.grads.flat, op=dist.ReduceOp.SUM, group=self.group)
        elif world > 1:
            # bucketed all-reduce overlapped with backward: the head bucket flies during the layers' wgrad kernels,
            # layer l's bucket during the wgrad of layer l-1 (the weight gradients are batched over all T steps, so
            # they only exist once BPTT is over: SURVEY.md section 8e)
            pending = []
            plan.backward(dpred, out=self._grad_views(),
                          on_ready=lambda i: pending.append(self.grads.all_reduce_bucket_async(i)))
            for work in pending:
                work.wait()
        else:
            plan.backward(dpred, out=self._grad_views())            # train.py:109, written into the flat buffer
        self.optimizer.step(grad_scale=1.0 / world)                 # train.py:110 (mean over ranks folded in)
        return loss[0]

    def _step_native(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        model = self.model
        if getattr(model, "return_sequence", False):
            raise NotImplementedError("the native step trains on the last-step prediction (train.py:96-105)")
        plan = model.plan_for(x, True)
        pred, _ = plan.forward(x)                                   # train.py:96
        return self._loss_and_update(plan, pred, y, None, 0)

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """One optimizer step on this rank's shard; returns the (un-synchronised) local loss tensor."""
        if self.native:
            return self._step_native(x, y)
        self.grads.zero()                                   # train.py:108
        pred = self.model(x.float() if x.dtype != torch.float32 and self.model.precision != "bf16" else x)   # train.py:96
        loss = training_loss(pred, y, self.crop)            # train.py:102,105
        loss.backward()                                     # train.py:109
        self.grads.all_reduce_mean()
        self.optimizer.step()                               # train.py:110
        return loss.detach()

    def step_windows(self, bank, win_start: torch.Tensor, seq_len: int) -> torch.Tensor:
        """One optimizer step on windows of a frame bank: sample b = bank.frames[i_b : i_b + T] with target
        bank.targets[i_b + T - 1] (dataset.py:600-601).  `win_start`: int32 [B] on the device (what HostFeeder delivers
        from the loader's index batch) or on the host."""
        if bank.targets is None:
            raise ValueError("step_windows needs a FrameBank with targets")
        model = self.model
        starts = win_start if (win_start.is_cuda and win_start.dtype == torch.int32) else \
            win_start.to(device=bank.frames.device, dtype=torch.int32)
        bank.check_windows(win_start, seq_len)
        if not self.native:
            self.grads.zero()
            pred = model.forward_windows(bank, starts, seq_len)
            y = bank.targets.index_select(0, starts.long() + (seq_len - 1))
            loss = training_loss(pred, y, self.crop)
            loss.backward()
            self.grads.all_reduce_mean()
            self.optimizer.step()
            return loss.detach()
        if getattr(model, "return_sequence", False):
            raise NotImplementedError("the native step trains on the last-step prediction (train.py:96-105)")
        _, H, W, _ = bank.frames.shape
        plan = model.plan_for_shape(starts.shape[0], int(seq_len), H, W, bank.frames.device, True)
        bank.check_plan(plan)
        pred, _ = plan.forward_bank(bank.frames, starts.contiguous())
        return self._loss_and_update(plan, pred, bank.targets, starts, seq_len - 1)

    # ---- the native step as one CUDA graph
    def capture(self, x: torch.Tensor = None, y: torch.Tensor = None, bank=None, win_start: torch.Tensor = None,
                seq_len: int = None, warmup: int = 2):
        """Records one native training step into a CUDA graph over STATIC input buffers (`self.static_inputs`): either
        (x, y) device tensors or (bank, win_start).  Afterwards `replay()` runs the step with whatever the caller
        copied into those buffers; nothing on the host changes between replays (Adam's step count and learning rate
        live in device memory).  Single-process only: NCCL collectives are left out of the graph."""
        if not self.native:
            raise RuntimeError("graph capture is for the native step")
        if self._world() > 1:
            raise RuntimeError("capture() is single-process; with several ranks run step() (the all-reduce stays eager)")
        if bank is not None:
            static = (win_start.to(device=bank.frames.device, dtype=torch.int32).clone(),)
            run = lambda: self.step_windows(bank, static[0], seq_len)
        else:
            static = (x.clone(), y.clone())
            run = lambda: self.step(static[0], static[1])
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):       # plans, tensor maps and kernel attributes are set up outside the capture
                run()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = run()
        self._graph, self._graph_loss, self.static_inputs = graph, loss, static
        return self.static_inputs

    def replay(self) -> torch.Tensor:
        """Runs the captured step on the current contents of `static_inputs`; returns the (static) loss tensor."""
        if self._graph is None:
            raise RuntimeError("replay() before capture()")
        self._graph.replay()
        self.optimizer.step_count += 1
        return self._graph_loss
