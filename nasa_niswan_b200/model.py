"""Drop-in replacements for the reference's `model.py` ConvLSTM classes (model.py:196-274).

Same constructor signatures, parameter names/shapes (`layers.{l}.conv.weight|bias`,
`conv.weight|bias`; checkpoint contract utils.py:27,39), default initialisation and forward
contracts, so `train.py`, `utils.py` and the notebooks can `from model import ConvLSTM`
unchanged.  The arithmetic runs in libnint.so (hand-written sm_100a kernels behind the C ABI of
include/nint.h); there is no CPU or eager-PyTorch fallback.
"""
import weakref
from collections import OrderedDict
from typing import Sequence

import torch
import torch.nn as nn

from . import ops
from .engine import Plan

_MAX_CACHED_PLANS = 4
_MAX_CELL_PLANS = 256


def _deterministic_default() -> bool:
    """The reference pins run-to-run reproducibility with `seed()` (utils.py:77-88: cudnn.deterministic = True); the
    same switch (or torch.use_deterministic_algorithms) selects the fixed-order gradient reductions here."""
    return bool(torch.backends.cudnn.deterministic or torch.are_deterministic_algorithms_enabled())


class _PlanCache:
    """LRU of plans (each owns a device workspace and a ctypes handle).  Plans are derived state: copying or pickling
    the owning module (copy.deepcopy, torch.save(model), swa_utils.AveragedModel) yields an EMPTY cache."""

    def __init__(self):
        self._plans = OrderedDict()

    def get(self, key, factory):
        p = self._plans.get(key)
        if p is None:
            p = factory()
            self._plans[key] = p
            while len(self._plans) > _MAX_CACHED_PLANS:
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        return p

    def clear(self):
        self._plans.clear()

    def __deepcopy__(self, memo):
        return type(self)()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()


class _CellPlanPool(_PlanCache):
    """Training plans of a standalone cell: every step that awaits its backward owns one (its saved gates live in the
    plan's workspace), so a user loop `for t: h, c = cell(x[t], (h, c))` followed by one `backward()` works like the
    reference's autograd module.  A plan returns to the pool when its autograd node dies."""

    def __init__(self):
        super().__init__()
        self._free = {}
        self._all = []      # strong references: the op registry and the autograd nodes only hold plan ids

    def acquire(self, key, factory):
        free = self._free.setdefault(key, [])
        if free:
            return free.pop()
        if len(self._all) >= _MAX_CELL_PLANS:
            raise RuntimeError(f"more than {_MAX_CELL_PLANS} ConvLSTMCell steps are awaiting backward()")
        plan = factory()
        self._all.append(plan)
        ref = weakref.ref(plan)
        plan.release = lambda f=free, r=ref: f.append(r()) if r() is not None else None
        return plan

    def clear(self):
        super().clear()
        self._free.clear()
        self._all.clear()


class ConvLSTMCell(nn.Module):
    """model.py:196-231.  `forward(x, (h, c)) -> (h, c)`: one fused step on the GPU, differentiable w.r.t. x, h, c and
    the parameters (the step's backward is one fused gate-backward launch, two dgrad launches and one wgrad launch)."""

    def __init__(self, input_channels, hidden_channels, kernel_size, bias=True, precision="bf16"):
        super().__init__()
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.kernel_size = kernel_size
        self.padding = kernel_size // 2
        self.bias = bias
        self.precision = precision
        # parameter container only (never called): same names, shapes and default init as the reference
        self.conv = nn.Conv2d(in_channels=self.input_channels + self.hidden_channels,
                              out_channels=4 * self.hidden_channels, kernel_size=self.kernel_size,
                              padding=self.padding, bias=self.bias)
        self._plans = _CellPlanPool()

    def forward(self, x, hidden_state):
        h, c = hidden_state
        if not x.is_cuda:
            raise RuntimeError("ConvLSTMCell runs on CUDA (B200) only: there is no CPU fallback")
        training = torch.is_grad_enabled() and (x.requires_grad or h.requires_grad or c.requires_grad or
                                                any(p.requires_grad for p in self.parameters()))
        B, _, H, W = x.shape
        key = (B, H, W, self.precision, x.device, training, _deterministic_default())
        if training:
            plan = self._plans.acquire(key, lambda: self._make_plan(B, H, W, x.device, True))
        else:
            plan = self._plans.get(key, lambda: self._make_plan(B, H, W, x.device, False))
        return torch.ops.nint.cell_forward(x, h, c, self.conv.weight, self.conv.bias, ops.register_plan(plan))

    def _make_plan(self, B, H, W, device, training):
        plan = Plan(B, 1, H, W, self.input_channels, [self.hidden_channels], [self.kernel_size],
                    precision=self.precision, training=training, device=device, input_grad=training,
                    deterministic=training and _deterministic_default())
        hw = torch.zeros(1, self.hidden_channels, 1, 1, device=device)
        plan.set_head(hw, torch.zeros(1, device=device))
        return plan


class ConvLSTM(nn.Module):
    """model.py:234-274.  `forward(x[B,T,C,H,W]) -> [B,1,H,W]` (head on the last layer's h at the
    last step).  `return_sequence=True` restates the commented-out variant (model.py:264,272,274)
    that test.ipynb:273 was run against: returns `(pred, hs[B,T,H,W])`.
    `precision`: "bf16" (bf16 operands, fp32 accumulate/state) or "tf32".
    x may be fp32 or (bf16 precision) bf16; `forward_windows(bank, starts)` reads the windows from an HBM-resident
    `preprocess.FrameBank` instead (dataset.py:551-637)."""

    def __init__(self, input_channels, hidden_channels: Sequence[int], kernel_size: Sequence[int], num_layers,
                 precision: str = "bf16", return_sequence: bool = False):
        super().__init__()
        assert len(hidden_channels) == num_layers, "The length of hidden_channels should be equal to num_layers"
        assert len(kernel_size) == num_layers, "The length of kernel_size should be equal to num_layers"
        self.num_layers = num_layers
        self.input_channels = input_channels
        self.precision = precision
        self.return_sequence = return_sequence
        layers = []
        for i in range(self.num_layers):
            in_ch = input_channels if i == 0 else hidden_channels[i - 1]
            layers.append(ConvLSTMCell(in_ch, hidden_channels[i], kernel_size[i], precision=precision))
        self.layers = nn.ModuleList(layers)
        self.conv = nn.Conv2d(hidden_channels[-1], 1, 1)
        self._plans = _PlanCache()

    def _params(self):
        ps = []
        for cell in self.layers:
            ps += [cell.conv.weight, cell.conv.bias]
        return ps + [self.conv.weight, self.conv.bias]

    def plan_for_shape(self, B, T, H, W, device, training, set_params=True, input_grad=False):
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        det = bool(training) and _deterministic_default()
        key = (B, T, H, W, bool(training), self.return_sequence, self.precision, device, bool(input_grad), det)
        hidden = [c.hidden_channels for c in self.layers]
        ks = [c.kernel_size for c in self.layers]
        plan = self._plans.get(key, lambda: Plan(B, T, H, W, self.input_channels, hidden, ks, precision=self.precision,
                                                 training=training, return_sequence=self.return_sequence,
                                                 device=device, deterministic=det, input_grad=input_grad))
        if set_params:
            for l, cell in enumerate(self.layers):
                plan.set_weights(l, cell.conv.weight, cell.conv.bias)
            plan.set_head(self.conv.weight, self.conv.bias)
        return plan

    def plan_for(self, x, training, set_params=True, input_grad=False):
        B, T, C, H, W = x.size()
        if C != self.input_channels:
            raise ValueError(f"expected {self.input_channels} input channels, got {C}")
        return self.plan_for_shape(B, T, H, W, x.device, training, set_params, input_grad)

    def _wants_grad(self, params):
        return torch.is_grad_enabled() and any(p.requires_grad for p in params)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("ConvLSTM runs on CUDA (B200) only: there is no CPU fallback; move the module and "
                               "its input to the GPU")
        params = self._params()
        need_dx = torch.is_grad_enabled() and x.requires_grad
        training = need_dx or self._wants_grad(params)
        plan = self.plan_for(x, training, set_params=False, input_grad=need_dx)   # the op repacks the weights it is handed
        # torch.library op (ops.py): forward = nint_forward (T x L fused cell steps + head), autograd = nint_backward
        pred, seq = torch.ops.nint.convlstm_forward(x, params, ops.register_plan(plan))
        return (pred, seq) if self.return_sequence else pred

    def forward_windows(self, bank, win_start, seq_len):
        """bank: preprocess.FrameBank (frames resident in HBM in the operand layout); win_start: int tensor [B] of
        window start frames -> the same outputs as forward(x) with x[b] = frames[win_start[b] : win_start[b]+T]."""
        params = self._params()
        starts = win_start.to(device=bank.frames.device, dtype=torch.int32).contiguous()
        bank.check_windows(win_start, seq_len)
        N, H, W, _ = bank.frames.shape
        plan = self.plan_for_shape(starts.shape[0], int(seq_len), H, W, bank.frames.device, self._wants_grad(params),
                                   set_params=False)
        bank.check_plan(plan)
        pred, seq = torch.ops.nint.convlstm_forward_bank(bank.frames, starts, params, ops.register_plan(plan))
        return (pred, seq) if self.return_sequence else pred

    def release_workspaces(self):
        self._plans.clear()
        for cell in self.layers:
            cell._plans.clear()
