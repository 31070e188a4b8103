"""GPU (B200): the input path (bf16 windows, HBM-resident frame bank, fused preprocessing), the API surface added in
round 2 (differentiable cell, input gradient, arbitrary hidden sizes, deterministic reductions, CUDA-graph training
step) and parity against the CPU oracle at the BASELINE.json sizes.  Metric and bars as in test_gpu_parity.py:
max|a-b| / max|b| vs the fp32 oracle, <= 1e-3 (tf32) / <= 2e-2 (bf16)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import convlstm_oracle as O

pytestmark = pytest.mark.gpu

TOL = {"tf32": 1e-3, "bf16": 2e-2}


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = False


def _fresh(C, hidden, ks, precision, seed=0, **kw):
    """(net on the GPU, CPU copies of its parameters)"""
    from nasa_niswan_b200 import ConvLSTM
    torch.manual_seed(seed)
    net = ConvLSTM(C, hidden, ks, len(hidden), precision=precision, **kw)
    params = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return net.cuda(), params


def _oracle_shared_dpred(x, y, params, L, crop=None):
    """oracle forward + backward of d(MSE+L1)/dpred taken at the ORACLE's prediction (see test_gpu_parity.py:
    L1's sign() is discontinuous, so both sides back-propagate the same upstream gradient)"""
    leaf = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    pred = O.convlstm_forward(x, leaf, L)
    dpred = torch.autograd.grad(O.training_loss(pred, y, crop), pred, retain_graph=True)[0]
    pred.backward(dpred)
    return pred.detach(), dpred, {k: v.grad for k, v in leaf.items()}


def _check_grads(net, ref_grads, tol):
    worst = 0.0
    for k, p in net.named_parameters():
        e = O.max_abs_normalised(p.grad.cpu(), ref_grads[k])
        worst = max(worst, e)
        assert e < tol, (k, e)
    return worst


def _round_tf32(t):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, on the 13 dropped mantissa bits"""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


# ------------------------------------------------------------------------------------------------ input path
def test_bf16_input_is_bit_identical_to_fp32_input():
    """host-staged bf16 windows (half the PCIe bytes of train.py:92) give exactly the fp32 path's results: the fp32
    path rounds x to bf16 while packing it"""
    net, _ = _fresh(21, [64], [3], "bf16")
    x = torch.randn(3, 4, 21, 30, 40, device="cuda")
    with torch.no_grad():
        a = net(x)
        b = net(x.to(torch.bfloat16))
    assert torch.equal(a, b)
    net32, _ = _fresh(21, [64], [3], "tf32")
    with pytest.raises(TypeError, match="bf16"):
        net32(x.to(torch.bfloat16))


@pytest.mark.parametrize("cfg", [("bf16", [64], [3]), ("bf16", [32], [5]), ("tf32", [64], [3]), ("bf16", [64, 32], [3, 3])],
                         ids=["bf16_pair_wgrad", "bf16_single_cta_wgrad", "tf32", "two_layers"])
def test_frame_bank_windows_match_explicit_windows(cfg):
    """nint_forward_bank: windows cut by TMA coordinates (frame = start[b] + t) out of an HBM-resident bank give the
    same forward (bit for bit) and the same gradients as the explicit [B,T,C,H,W] tensor of those windows
    (dataset.py:551-637: sliding_window_view over the record)."""
    from nasa_niswan_b200.preprocess import FrameBank
    precision, hidden, ks = cfg
    C, H, W, T, N = 21, 26, 40, 5, 17
    net, _ = _fresh(C, hidden, ks, precision)
    torch.manual_seed(2)
    record = torch.randn(N, C, H, W, device="cuda")
    starts = torch.tensor([3, 0, 12, 7, 7], dtype=torch.int64)
    bank = FrameBank.from_frames(record, precision)
    windows = torch.stack([record[s:s + T] for s in starts.tolist()])          # [B,T,C,H,W]
    prev = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True      # fixed-order reductions: the two paths must agree bit for bit
    try:
        pred_w = net(windows)
        g = torch.randn_like(pred_w)
        pred_w.backward(g)
        grads_w = [p.grad.clone() for p in net.parameters()]
        net.zero_grad(set_to_none=True)
        pred_b = net.forward_windows(bank, starts, T)
        pred_b.backward(g)
    finally:
        torch.backends.cudnn.deterministic = prev
    assert torch.equal(pred_w, pred_b)
    for a, p in zip(grads_w, net.parameters()):
        assert torch.equal(a, p.grad)
    with pytest.raises(IndexError):
        net.forward_windows(bank, torch.tensor([N - T + 1]), T)
    # indices that only exist on the device are not checked on the host: frames outside the bank read as zeros through
    # TMA (inputs) and as zero targets (loss) -- never a fault
    with torch.no_grad():
        wild = net.forward_windows(bank, torch.tensor([N + 5, -3, 0, 1, 2], dtype=torch.int32, device="cuda"), T)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(wild).all())
    assert torch.equal(wild[2], pred_b[1].detach())          # window start 0 again: per-sample arithmetic, bit for bit


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
@pytest.mark.parametrize("mode", ["reflect", "reference_rnn"])
@pytest.mark.parametrize("shape", [(3, 20, 90, 144, 100, 154, 0), (2, 4, 12, 20, 22, 30, 3), (1, 0, 9, 7, 9, 7, 0)])
def test_fused_preprocessing_writes_the_operand_layout(shape, mode, precision):
    """nint_fuse_inputs_bank (stack + z-score + statics + geophysical halo -> channels-last E with the ones lane, ONE
    kernel) is bit-exact against the oracle's fp32 pipeline (dataset.py:520-537, 67-98) rounded the way the model's
    input packing rounds it."""
    from nasa_niswan_b200.preprocess import FrameBank, input_layout
    N, L, H, W, Hp, Wp, S = shape
    rng = np.random.default_rng(7)
    lev = rng.standard_normal((N, L, H, W)).astype(np.float32) * 3 + 1
    emis = np.abs(rng.standard_normal((N, H, W))).astype(np.float32)
    mean = rng.standard_normal(L + 1).astype(np.float32)
    std = (rng.random(L + 1) + 0.5).astype(np.float32)
    statics = O.normalise_static_attributes(rng.standard_normal((S, H, W))).astype(np.float32) if S else None
    want = torch.from_numpy(O.fuse_inputs(lev, emis, mean, std, (Hp, Wp), mode, statics))     # [N, C, Hp, Wp] fp32
    cu = lambda a: None if a is None else torch.from_numpy(a).cuda()
    bank = FrameBank.from_fields(cu(lev), cu(emis), cu(mean), cu(std), (Hp, Wp), mode, cu(statics), precision)
    C = L + 1 + S
    c_pad, ones = input_layout(C)
    assert tuple(bank.frames.shape) == (N, Hp, Wp, c_pad)
    got = bank.frames.cpu()
    ref = want.permute(0, 2, 3, 1).contiguous()
    ref = ref.to(torch.bfloat16) if precision == "bf16" else _round_tf32(ref)
    assert torch.equal(got[..., :C], ref)
    if ones >= 0:
        assert bool((got[..., ones].float() == 1).all()) and bool((got[..., ones + 1:].float() == 0).all())
    # and the two-pass path (fuse_inputs -> fp32 NCHW, then pack) lands on the same bits
    two = FrameBank.from_frames(want.cuda(), precision)
    assert torch.equal(two.frames, bank.frames)


def test_trainer_step_windows_matches_step_on_explicit_windows():
    """Trainer.step_windows (indices into the bank; targets read through the same indices by the fused loss kernel)
    == Trainer.step on the gathered windows, step for step"""
    from nasa_niswan_b200.parallel import Trainer
    from nasa_niswan_b200.preprocess import FrameBank
    C, H, W, T, N, B = 21, 20, 24, 4, 30, 6
    torch.manual_seed(4)
    record, targets = torch.randn(N, C, H, W, device="cuda"), torch.randn(N, H, W, device="cuda")
    bank = FrameBank.from_frames(record, "bf16", targets=targets)
    net_a, _ = _fresh(C, [32], [3], "bf16", seed=9)
    net_b, _ = _fresh(C, [32], [3], "bf16", seed=9)
    ta, tb = Trainer(net_a, lr=1e-3), Trainer(net_b, lr=1e-3)
    gen = torch.Generator().manual_seed(0)
    prev = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True      # fixed-order gradient sums: the two input paths agree bit for bit, and
    try:                                           # Adam cannot amplify a last-bit difference of a near-zero gradient
        for _ in range(3):
            starts = torch.randint(0, N - T + 1, (B,), generator=gen)
            x = torch.stack([record[s:s + T] for s in starts.tolist()])
            y = targets[starts + T - 1]
            la = ta.step(x, y)
            lb = tb.step_windows(bank, starts.to(torch.int32).cuda(), T)
            assert abs(float(la) - float(lb)) <= 1e-6 * max(1.0, abs(float(la)))
    finally:
        torch.backends.cudnn.deterministic = prev
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        assert torch.equal(pa, pb)


def test_host_feeder_carries_bf16_windows_and_index_batches():
    from nasa_niswan_b200.parallel import HostFeeder
    feeder = HostFeeder("cuda:0")
    xh = torch.randn(2, 3, 5, 8, 8).to(torch.bfloat16).pin_memory()
    ih = torch.arange(4, dtype=torch.int32).pin_memory()
    for _ in range(3):
        slot = feeder.put(xh, ih)
        xd, idx = feeder.get(slot)
        assert xd.dtype == torch.bfloat16 and idx.dtype == torch.int32
        torch.cuda.current_stream().synchronize()
        assert torch.equal(xd.cpu(), xh) and torch.equal(idx.cpu(), ih)
        feeder.release(slot)


# ------------------------------------------------------------------------------------------------ advisor findings
def test_second_backward_of_one_forward_is_refused():
    """BPTT overwrites the saved gates with their gradients in place; a second backward over the same forward
    (retain_graph=True, or two autograd.grad calls) must raise instead of silently differentiating the dgates"""
    net, _ = _fresh(5, [16], [3], "bf16")
    x = torch.randn(2, 3, 5, 12, 16, device="cuda")
    pred = net(x)
    pred.sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="already back-propagated|consumed"):
        pred.sum().backward()
    # the C ABI refuses it too (engine-level callers)
    plan = net.plan_for(x, True)
    plan.forward(x)
    plan.backward(torch.ones(2, 1, 12, 16, device="cuda"))
    with pytest.raises(RuntimeError, match="consumed"):
        plan.backward(torch.ones(2, 1, 12, 16, device="cuda"))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_a_non_current_device():
    """a model on cuda:1 while the current device is 0 must launch on GPU 1 (device guard), and tensors on
    the wrong device are rejected"""
    assert torch.cuda.current_device() == 0
    from nasa_niswan_b200 import ConvLSTM
    torch.manual_seed(0)
    net = ConvLSTM(5, [16], [3], 1, precision="tf32")
    params = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.to("cuda:1")
    x = torch.randn(2, 3, 5, 12, 16)
    pred = net(x.to("cuda:1"))
    pred.sum().backward()
    ref = O.convlstm_forward(x, params, 1)
    assert O.max_abs_normalised(pred.detach().cpu(), ref) < TOL["tf32"]
    with pytest.raises(RuntimeError, match="lives on"):
        net.plan_for(x.to("cuda:1"), False).forward(x.to("cuda:0"))


def _fused_tail_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from nasa_niswan_b200 import ConvLSTM
    from nasa_niswan_b200.parallel import Trainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.backends.cudnn.deterministic = True          # fixed-order local gradients: the two tails see the same inputs
    C, H, W, T, B = 21, 20, 24, 4, 4
    results = {}
    for fused in ("1", "0"):
        os.environ["NINT_DP_FUSED"] = fused
        torch.manual_seed(5)
        net = ConvLSTM(C, [32], [3], 1, precision="bf16").to(dev)
        tr = Trainer(net, lr=1e-3, scheduler_config=(2, 0.5))
        assert (tr.sym is not None) == (fused == "1"), "fused NVLink tail did not come up"
        gen = torch.Generator().manual_seed(100 + rank)          # every rank trains on its own shard
        losses = []
        for step in range(5):
            x = torch.randn(B, T, C, H, W, generator=gen).to(dev)
            y = torch.randn(B, H, W, generator=gen).to(dev)
            losses.append(float(tr.step(x, y)))
            if step % 2 == 1:
                tr.end_epoch()
        results[fused] = ([p.detach().cpu() for p in net.parameters()], losses)
    torch.save(results, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with NVLink peer access")
def test_fused_nvlink_step_tail_matches_nccl(tmp_path):
    """nint_dp_allreduce_adam (cross-rank barrier + gradient sum over NVLink peer memory + Adam in one kernel) against
    ncclAllReduce + the Adam kernel: same parameters on every rank after five steps on rank-specific shards, and the
    replicas stay bit-identical to each other"""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    world = 2
    mp.spawn(_fused_tail_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    for fused in ("1", "0"):
        for a, b in zip(res[0][fused][0], res[1][fused][0]):
            assert torch.equal(a, b), "replicas diverged"
    for a, b in zip(res[0]["1"][0], res[0]["0"][0]):
        assert O.max_abs_normalised(a, b) < 1e-6
    assert res[0]["1"][1] != res[1]["1"][1]                       # different shards -> different local losses
    for la, lb in zip(res[0]["1"][1], res[0]["0"][1]):
        assert abs(la - lb) < 1e-5 * max(1.0, abs(la))


# ------------------------------------------------------------------------------------------------ API completeness
@pytest.mark.parametrize("cfg", [("tf32", 5, 32, 3), ("bf16", 21, 64, 5), ("tf32", 3, 10, 3)], ids=["tf32_h32", "bf16_h64_k5", "tf32_h10"])
def test_cell_is_differentiable_like_the_reference_module(cfg):
    """model.py:216-231 is an ordinary autograd module: gradients w.r.t. x, h, c, weight and bias of one step, and a
    user-written time loop over the cell followed by one backward()"""
    from nasa_niswan_b200 import ConvLSTMCell
    precision, C, hc, k = cfg
    tol = TOL[precision]
    torch.manual_seed(1)
    cell = ConvLSTMCell(C, hc, k, precision=precision)
    w, b = cell.conv.weight.detach().clone(), cell.conv.bias.detach().clone()
    cell = cell.cuda()
    B, H, W = 2, 14, 19
    x, h, c = torch.randn(B, C, H, W), torch.randn(B, hc, H, W) * 0.5, torch.randn(B, hc, H, W) * 0.5
    gh, gc = torch.randn(B, hc, H, W), torch.randn(B, hc, H, W)
    leaves = [t.clone().requires_grad_(True) for t in (x, h, c, w, b)]
    rh, rc = O.cell_forward(*leaves)
    torch.autograd.backward([rh, rc], [gh, gc])
    dev = [t.clone().cuda().requires_grad_(True) for t in (x, h, c)]
    oh, oc = cell(dev[0], (dev[1], dev[2]))
    assert O.max_abs_normalised(oh.detach().cpu(), rh.detach()) < tol and O.max_abs_normalised(oc.detach().cpu(), rc.detach()) < tol
    torch.autograd.backward([oh, oc], [gh.cuda(), gc.cuda()])
    got = [dev[0].grad, dev[1].grad, dev[2].grad, cell.conv.weight.grad, cell.conv.bias.grad]
    for name, a, r in zip(("dx", "dh", "dc", "dw", "db"), got, leaves):
        assert O.max_abs_normalised(a.cpu(), r.grad) < tol, name
    # a user loop: three steps, one backward (every step keeps its own saved gates until its backward ran)
    cell.zero_grad(set_to_none=True)
    xs = torch.randn(3, B, C, H, W)
    lw, lb = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    rh, rc = torch.zeros(B, hc, H, W), torch.zeros(B, hc, H, W)
    dh, dc = torch.zeros(B, hc, H, W, device="cuda"), torch.zeros(B, hc, H, W, device="cuda")
    for t in range(3):
        rh, rc = O.cell_forward(xs[t], rh, rc, lw, lb)
        dh, dc = cell(xs[t].cuda(), (dh, dc))
    (rh * gh).sum().backward()
    (dh * gh.cuda()).sum().backward()
    assert O.max_abs_normalised(cell.conv.weight.grad.cpu(), lw.grad) < tol
    assert O.max_abs_normalised(cell.conv.bias.grad.cpu(), lb.grad) < tol
    with torch.no_grad():                              # and the forward-only path still works
        oh2, _ = cell(dev[0].detach(), (dev[1].detach(), dev[2].detach()))
    assert torch.equal(oh2, oh.detach())


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_input_gradient_matches_autograd(precision):
    """x.grad through model.py:219-220 (the reference's autograd yields it; e.g. for saliency maps)"""
    net, params = _fresh(21, [64, 32], [3, 3], precision)
    x, y = torch.randn(2, 3, 21, 18, 22), torch.randn(2, 18, 22)
    xr = x.clone().requires_grad_(True)
    ref_pred = O.convlstm_forward(xr, params, 2)
    g = torch.autograd.grad(O.training_loss(ref_pred, y), ref_pred, retain_graph=True)[0]   # shared upstream gradient
    ref_pred.backward(g)
    xd = x.clone().cuda().requires_grad_(True)
    pred = net(xd)
    pred.backward(g.cuda())
    assert xd.grad.shape == x.shape
    assert O.max_abs_normalised(xd.grad.cpu(), xr.grad) < TOL[precision]
    # parameters still get their gradients on this path
    leaf = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    O.convlstm_forward(x, leaf, 2).backward(g)
    _check_grads(net, {k: v.grad for k, v in leaf.items()}, TOL[precision])


@pytest.mark.parametrize("cfg", [(5, [10], [3]), (21, [70], [3]), (5, [24, 5], [3, 5]), (8, [100, 48], [3, 3])],
                         ids=["h10", "h70", "h24_h5", "h100_h48"])
@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_arbitrary_hidden_sizes(cfg, precision):
    """model.py:207 accepts any channel counts (the notebooks build ConvLSTM(5, 10, 3, 2)-style models,
    dataset_config.ipynb:759-761): hidden sizes run padded to the kernels' granularity with zero weights, so outputs,
    gradients and state_dict shapes are those of the unpadded model"""
    C, hidden, ks = cfg
    net, params = _fresh(C, hidden, ks, precision)
    x, y = torch.randn(2, 3, C, 17, 21), torch.randn(2, 17, 21)
    ref_pred, dpred, ref_grads = _oracle_shared_dpred(x, y, params, len(hidden))
    pred = net(x.cuda())
    assert O.max_abs_normalised(pred.detach().cpu(), ref_pred) < TOL[precision]
    pred.backward(dpred.cuda())
    _check_grads(net, ref_grads, TOL[precision])
    assert [tuple(p.shape) for p in net.parameters()] == [tuple(v.shape) for v in params.values()]


def test_deterministic_mode_is_bit_reproducible():
    """utils.py:77-88 (`seed()`) pins cudnn.deterministic; the same switch makes every gradient reduction here run in
    a fixed order (per-split partial sums instead of fp32 atomics): two runs agree bit for bit, and with the default
    (atomic) mode to rounding"""
    from nasa_niswan_b200.utils import seed
    x = torch.randn(4, 5, 21, 40, 48, device="cuda")
    g = torch.randn(4, 1, 40, 48, device="cuda")

    def run():
        net, _ = _fresh(21, [64, 32], [3, 3], "bf16", seed=3)
        net(x).backward(g)
        return [p.grad.clone() for p in net.parameters()]
    prev = torch.backends.cudnn.deterministic
    try:
        seed(0)
        a, b = run(), run()
    finally:
        torch.backends.cudnn.deterministic = prev
    c = run()
    for ga, gb, gc in zip(a, b, c):
        assert torch.equal(ga, gb)
        assert O.max_abs_normalised(gc.cpu(), ga.cpu()) < 1e-5


def test_sub_batch_major_schedule_changes_nothing(monkeypatch):
    """NINT_SUB_BATCH: images [b0, b0+n) run all T steps before the next slice (L2 residency of the recurrent
    operands); per-tile arithmetic is untouched, so forward and (deterministic) gradients are bit-identical"""
    x = torch.randn(6, 4, 21, 30, 40, device="cuda")
    g = torch.randn(6, 1, 30, 40, device="cuda")

    def run(sub):
        if sub:
            monkeypatch.setenv("NINT_SUB_BATCH", str(sub))
        else:
            monkeypatch.delenv("NINT_SUB_BATCH", raising=False)
        net, _ = _fresh(21, [64, 32], [3, 3], "bf16", seed=3)
        pred = net(x)
        pred.backward(g)
        with torch.no_grad():
            inf = net(x)
        return pred.detach(), inf, [p.grad.clone() for p in net.parameters()]
    prev = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        ref = run(0)
        for sub in (2, 4):
            got = run(sub)
            assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1])
            for a, b in zip(ref[2], got[2]):
                assert torch.equal(a, b)
    finally:
        torch.backends.cudnn.deterministic = prev


@pytest.mark.parametrize("shape", [(9, 6, 40, 36), (40, 4, 40, 36)])
@pytest.mark.parametrize("hidden,ks,precision,seq", [([64], [3], "bf16", False), ([64, 32], [3, 5], "bf16", False),
                                                     ([32], [3], "tf32", False), ([16, 16, 16], [3, 3, 3], "bf16", True)])
def test_time_fused_launches_change_nothing(monkeypatch, hidden, ks, precision, seq, shape):
    """NINT_FUSE_STEPS (default on): all time steps of a layer run as ONE persistent launch whose tiles wait for the
    previous step's tiles of the same image instead of for the whole previous launch.  Per-tile arithmetic is untouched,
    so training forward, inference forward (2-slot h ring) and the deterministic gradients are bit-identical to the
    one-launch-per-step schedule.  9 images of 3 x 5 tiles: fewer tile groups per step than CTA pairs, every cluster
    jumps a step between groups; 40 images: several groups per cluster and step"""
    B, T = shape[:2]
    x = torch.randn(B, T, 21, shape[2], shape[3], device="cuda")

    from nasa_niswan_b200 import ConvLSTM

    def run(fuse):
        monkeypatch.setenv("NINT_FUSE_STEPS", "3" if fuse else "0")
        torch.manual_seed(11)
        net = ConvLSTM(21, hidden, ks, len(hidden), precision=precision, return_sequence=seq).cuda()
        if seq:
            pred, hs = net(x)
            (pred.sum() + (hs * torch.linspace(0.5, 1.5, T, device="cuda").view(1, T, 1, 1)).sum()).backward()
        else:
            pred = net(x)
            pred.backward(torch.ones_like(pred) * 0.25)
        with torch.no_grad():
            inf = net(x)
            inf = inf[0] if seq else inf
        return pred.detach(), inf, [p.grad.clone() for p in net.parameters()]
    prev = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        ref = run(False)
        for _ in range(2):
            got = run(True)
            assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1])
            for a, b in zip(ref[2], got[2]):
                assert torch.equal(a, b)
    finally:
        torch.backends.cudnn.deterministic = prev


def test_time_fused_rule_follows_the_launch_length(monkeypatch):
    """default schedule: a layer's BPTT steps are ONE launch (after the step without a dgates_{t+1} operand) while a
    CTA pair walks at most 10 tile groups per step, one launch per step for long launches such as BASELINE cfg 2 at
    B=32; NINT_FUSE_STEPS=0 forbids fusing (counted through nint_launch_count, class 1 = dgrad + gate backward)"""
    from nasa_niswan_b200 import ConvLSTM, _lib
    lib = _lib.load()

    def bwd_launches(B, T, H, W, env):
        if env is None:
            monkeypatch.delenv("NINT_FUSE_STEPS", raising=False)
        else:
            monkeypatch.setenv("NINT_FUSE_STEPS", env)
        torch.manual_seed(0)
        net = ConvLSTM(21, [64], [3], 1, precision="bf16").cuda()
        pred = net(torch.randn(B, T, 21, H, W, device="cuda"))
        n0 = lib.nint_launch_count(1)
        pred.sum().backward()
        torch.cuda.synchronize()
        return lib.nint_launch_count(1) - n0
    assert bwd_launches(2, 6, 24, 24, None) == 2          # 2 images of 2 x 3 tiles: short launches, fused
    assert bwd_launches(2, 6, 24, 24, "0") == 6
    assert bwd_launches(32, 3, 90, 144, None) == 3        # 11.7 rounds of tile groups per CTA pair: one launch per step
    assert bwd_launches(32, 3, 90, 144, "2") == 2


def test_training_step_replays_as_one_cuda_graph():
    """the native step (forward, fused loss, BPTT, wgrad, Adam with its step count and lr in device memory) captured
    once and replayed: same parameters as the eager steps, including across a learning-rate change"""
    from nasa_niswan_b200.parallel import Trainer
    torch.backends.cudnn.deterministic = True     # fixed-order sums: eager and replayed steps see identical gradients
    C, H, W, T, B = 21, 20, 24, 4, 4
    torch.manual_seed(6)
    xs = [torch.randn(B, T, C, H, W, device="cuda") for _ in range(4)]
    ys = [torch.randn(B, H, W, device="cuda") for _ in range(4)]
    net_a, _ = _fresh(C, [32], [3], "bf16", seed=2)
    net_b, _ = _fresh(C, [32], [3], "bf16", seed=2)
    ta = Trainer(net_a, lr=1e-3, scheduler_config=(1, 0.5))
    tb = Trainer(net_b, lr=1e-3, scheduler_config=(1, 0.5))
    for t in (ta, tb):
        t.step(xs[0], ys[0])
    sx, sy = tb.capture(xs[0], ys[0], warmup=1)            # one eager warm-up step on the static buffers, then the capture
    ta.step(xs[0], ys[0])                                  # (capturing records the step; it does not run it)
    for i in range(1, 4):
        la = ta.step(xs[i], ys[i])
        sx.copy_(xs[i])
        sy.copy_(ys[i])
        lb = tb.replay()
        assert abs(float(la) - float(lb)) < 1e-5 * max(1.0, abs(float(la)))
        if i == 2:
            assert ta.end_epoch() == tb.end_epoch() == [5e-4]
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        assert O.max_abs_normalised(pb.detach().cpu(), pa.detach().cpu()) < 1e-4
    assert int(float(tb.optimizer.state[0])) == int(float(ta.optimizer.state[0]))
    torch.backends.cudnn.deterministic = False


# ------------------------------------------------------------------------------------------------ parity at size
@pytest.mark.parametrize("ksize", [3, 5])
def test_cfg2_full_batch_against_oracle(ksize):
    """BASELINE cfg 2 at its own size: B=32, T=12, 21 -> 64 channels, 90x144, bf16.  Tile walks, split-K shares and
    cluster counts depend on B, so this is the geometry the benchmark runs.  Forward vs the oracle; BPTT of a shared
    upstream gradient; then the whole chain loss -> gradients through the fused loss kernel (train.py:102-109) with the
    L1 sign taken at the GPU's own prediction."""
    B, T, C, H, W = 32, 12, 21, 90, 144
    net, params = _fresh(C, [64], [ksize], "bf16")
    torch.manual_seed(11)
    x, y = torch.randn(B, T, C, H, W), torch.randn(B, H, W)
    ref_pred, dpred, ref_grads = _oracle_shared_dpred(x, y, params, 1)
    xd = x.cuda()
    pred = net(xd)
    e_pred = O.max_abs_normalised(pred.detach().cpu(), ref_pred)
    assert e_pred < TOL["bf16"]
    pred.backward(dpred.cuda())
    worst = _check_grads(net, ref_grads, TOL["bf16"])
    # full chain: the fused loss kernel's gradient (sign taken at the GPU prediction) vs autograd of the oracle
    _, ref_loss, full_grads = O.forward_backward(x, y, params, 1)
    net.zero_grad(set_to_none=True)
    pred = net(xd)
    loss = F.mse_loss(pred.squeeze(1), y.cuda()) + F.l1_loss(pred.squeeze(1), y.cuda())
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 1e-3 * float(ref_loss)
    chain = _check_grads(net, full_grads, TOL["bf16"])
    print(f"cfg2 k{ksize} B=32: pred {e_pred:.2e}, grads (shared dpred) {worst:.2e}, grads (full chain) {chain:.2e}")


def test_cfg4_long_rollout_against_oracle_and_across_batch_sizes():
    """BASELINE cfg 4 at its own size: 120-step inference rollout on 90x144 with the state resident in HBM.  B=4 against
    the oracle; B=64 (the benchmarked batch) must reproduce the B=4 results bit for bit on the shared samples --
    per-sample arithmetic does not depend on the batch it rides in."""
    T, C, H, W = 120, 21, 90, 144
    net, params = _fresh(C, [64], [3], "bf16")
    torch.manual_seed(12)
    x4 = torch.randn(4, T, C, H, W)
    ref = O.convlstm_forward(x4, params, 1)
    with torch.no_grad():
        p4 = net(x4.cuda())
        e = O.max_abs_normalised(p4.cpu(), ref)
        assert e < TOL["bf16"]
        x64 = torch.randn(64, T, C, H, W, device="cuda")
        x64[:4] = x4.cuda()
        p64 = net(x64)
    assert torch.equal(p64[:4], p4)
    assert bool(torch.isfinite(p64).all())
    print(f"cfg4 T=120: pred error vs oracle {e:.2e}")


def test_cfg5_refined_grid_against_oracle():
    """BASELINE cfg 5 at its own grid: 3 layers x hidden 128, 5x5 kernels, 180x288, T=12 (B=1 so the CPU oracle finishes
    in about a minute)."""
    B, T, C, H, W = 1, 12, 21, 180, 288
    net, params = _fresh(C, [128, 128, 128], [5, 5, 5], "bf16")
    torch.manual_seed(13)
    x, y = torch.randn(B, T, C, H, W), torch.randn(B, H, W)
    ref_pred, dpred, ref_grads = _oracle_shared_dpred(x, y, params, 3)
    pred = net(x.cuda())
    e = O.max_abs_normalised(pred.detach().cpu(), ref_pred)
    assert e < TOL["bf16"]
    pred.backward(dpred.cuda())
    worst = _check_grads(net, ref_grads, TOL["bf16"])
    print(f"cfg5 180x288 3x128 k5: pred {e:.2e}, grads {worst:.2e}")


def test_shipped_model_pipeline_against_oracle():
    """The reference's one real recipe (launcher.sh:13-30): ConvLSTM(5, [64,32,16], [5,3,3]), T=48, 100x154 inputs built
    by the RNN dataset's halo (dataset.py:67-98 with its fliplr quirk), prediction cropped [5:95, 5:149]
    (train.py:102), MSE+L1.  Here: raw fields -> fused preprocessing kernel -> frame bank -> windows by index ->
    forward -> crop -> fused loss -> BPTT, against the oracle's fp32 pipeline."""
    from nasa_niswan_b200.preprocess import FrameBank
    T, L, H, W, Hp, Wp, N = 48, 4, 90, 144, 100, 154, 50
    crop = (5, 95, 5, 149)
    rng = np.random.default_rng(3)
    lev = rng.standard_normal((N, L, H, W)).astype(np.float32) * 2 + 0.5
    emis = np.abs(rng.standard_normal((N, H, W))).astype(np.float32)
    mean, std = rng.standard_normal(L + 1).astype(np.float32), (rng.random(L + 1) + 0.5).astype(np.float32)
    frames = torch.from_numpy(O.fuse_inputs(lev, emis, mean, std, (Hp, Wp), "reference_rnn"))     # [N,5,100,154]
    starts = [0, 2]
    x = torch.stack([frames[s:s + T] for s in starts])
    y = torch.randn(2, H, W)
    net, params = _fresh(5, [64, 32, 16], [5, 3, 3], "bf16")
    ref_pred, dpred, ref_grads = _oracle_shared_dpred(x, y, params, 3, crop)
    cu = lambda a: torch.from_numpy(a).cuda()
    bank = FrameBank.from_fields(cu(lev), cu(emis), cu(mean), cu(std), (Hp, Wp), "reference_rnn", None, "bf16")
    pred = net.forward_windows(bank, torch.tensor(starts), T)
    e = O.max_abs_normalised(pred.detach().cpu(), ref_pred)
    assert e < TOL["bf16"]
    pred.backward(dpred.cuda())
    worst = _check_grads(net, ref_grads, TOL["bf16"])
    print(f"shipped model T=48 100x154: pred {e:.2e}, grads {worst:.2e}")
