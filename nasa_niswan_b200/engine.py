"""Host-side plan object: one `nint_plan` (include/nint.h) plus the torch-allocated workspace it
is bound to.  PyTorch is plumbing here (device memory, streams); every byte of compute happens
in libnint.so."""
import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib

PRECISIONS = {"bf16": _lib.DTYPE_BF16, "tf32": _lib.DTYPE_TF32, "fp32": _lib.DTYPE_TF32}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_dev(t: torch.Tensor, name: str, shape=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the ConvLSTM hot path has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")


class Plan:
    """Geometry + workspace for `ConvLSTM(in_channels, hidden, ksize, L)` on inputs [B,T,C,H,W]."""

    def __init__(self, batch: int, seq_len: int, height: int, width: int, in_channels: int,
                 hidden: Sequence[int], ksize: Sequence[int], precision: str = "bf16", training: bool = False,
                 return_sequence: bool = False, device=None):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        if len(hidden) != len(ksize):
            raise ValueError("hidden and ksize must have the same length")
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the ConvLSTM hot path has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.B, self.T, self.H, self.W, self.C = batch, seq_len, height, width, in_channels
        self.hidden, self.ksize, self.L = list(hidden), list(ksize), len(hidden)
        self.training, self.return_sequence, self.precision = bool(training), bool(return_sequence), precision
        cfg = _lib.NintConfig()
        cfg.batch, cfg.seq_len, cfg.height, cfg.width = batch, seq_len, height, width
        cfg.in_channels, cfg.num_layers = in_channels, self.L
        for i, (h, k) in enumerate(zip(hidden, ksize)):
            cfg.hidden[i], cfg.ksize[i] = h, k
        cfg.dtype, cfg.training, cfg.return_sequence = PRECISIONS[precision], int(training), int(return_sequence)
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.nint_plan_create(ctypes.byref(cfg), ctypes.byref(self._h)), "nint_plan_create")
        self.workspace_bytes = self.lib.nint_plan_workspace_bytes(self._h)
        with torch.cuda.device(self.device):
            self.workspace = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.nint_plan_bind(self._h, _ptr(self.workspace), self.workspace_bytes, _stream()),
                       "nint_plan_bind")
        self.generation = 0          # bumped by every forward; backward must match
        self._weight_keys = [None] * (self.L + 1)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.nint_plan_destroy(h)

    # ---- parameters
    def set_weights(self, layer: int, weight: torch.Tensor, bias: Optional[torch.Tensor], force=True):
        """Repacks the fp32 OIHW masters into operand panels (two small kernels).  Done on EVERY call by default:
        tensor version counters are not a safe cache key -- torch's fused optimizers update parameters in place
        without bumping them, and a stale panel cache silently trains / evaluates with old weights."""
        key = (weight.data_ptr(), weight._version, None if bias is None else (bias.data_ptr(), bias._version))
        if not force and self._weight_keys[layer] == key:
            return
        hc = self.hidden[layer]
        cin = self.C if layer == 0 else self.hidden[layer - 1]
        _check_dev(weight, "weight", (4 * hc, cin + hc, self.ksize[layer], self.ksize[layer]))
        if bias is not None:
            _check_dev(bias, "bias", (4 * hc,))
        w = weight.detach().contiguous()
        b = None if bias is None else bias.detach().contiguous()
        _lib.check(self.lib.nint_plan_set_weights(self._h, layer, _ptr(w), _ptr(b), _stream()), "nint_plan_set_weights")
        self._weight_keys[layer] = key

    def set_head(self, weight: torch.Tensor, bias: torch.Tensor, force=True):
        key = (weight.data_ptr(), weight._version, bias.data_ptr(), bias._version)
        if not force and self._weight_keys[self.L] == key:
            return
        _check_dev(weight, "head weight", (1, self.hidden[-1], 1, 1))
        _check_dev(bias, "head bias", (1,))
        _lib.check(self.lib.nint_plan_set_head(self._h, _ptr(weight.detach().contiguous()),
                                               _ptr(bias.detach().contiguous()), _stream()), "nint_plan_set_head")
        self._weight_keys[self.L] = key

    # ---- state
    def reset_state(self):
        _lib.check(self.lib.nint_plan_reset_state(self._h, _stream()), "nint_plan_reset_state")

    def set_state(self, layer: int, h: torch.Tensor, c: torch.Tensor):
        shape = (self.B, self.hidden[layer], self.H, self.W)
        _check_dev(h, "h", shape)
        _check_dev(c, "c", shape)
        h, c = h.detach().contiguous(), c.detach().contiguous()
        _lib.check(self.lib.nint_plan_set_state(self._h, layer, _ptr(h), _ptr(c), _stream()), "nint_plan_set_state")

    def get_state(self, layer: int):
        shape = (self.B, self.hidden[layer], self.H, self.W)
        h = torch.empty(shape, dtype=torch.float32, device=self.device)
        c = torch.empty(shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.nint_plan_get_state(self._h, layer, _ptr(h), _ptr(c), _stream()), "nint_plan_get_state")
        return h, c

    # ---- compute
    def forward(self, x: torch.Tensor):
        _check_dev(x, "x", (self.B, self.T, self.C, self.H, self.W))
        x = x.detach().contiguous()
        pred = torch.empty((self.B, 1, self.H, self.W), dtype=torch.float32, device=self.device)
        seq = (torch.empty((self.B, self.T, self.H, self.W), dtype=torch.float32, device=self.device)
               if self.return_sequence else None)
        _lib.check(self.lib.nint_forward(self._h, _ptr(x), _ptr(pred), _ptr(seq), _stream()), "nint_forward")
        self.generation += 1
        return pred, seq

    def backward(self, dpred: Optional[torch.Tensor], dseq: Optional[torch.Tensor] = None, out=None, on_ready=None):
        """Returns ([grad_weight_l], [grad_bias_l], grad_head_weight, grad_head_bias).  `out`: optional
        (gw list, gb list, ghw, ghb) of preallocated contiguous fp32 tensors the gradients are written into
        (e.g. views of one flat all-reduce buffer).  `on_ready(i)`: called as soon as the kernels producing a
        gradient bucket are queued on the current stream -- i = L for the head, then i = L-1 .. 0 for the layers."""
        if dseq is not None:
            dseq = dseq.detach().contiguous().clone() if dpred is not None else dseq.detach().contiguous()
            _check_dev(dseq, "dseq", (self.B, self.T, self.H, self.W))
            if dpred is not None:   # pred is seq[:, -1]: fold its gradient in
                dseq[:, -1] += dpred.detach().reshape(self.B, self.H, self.W)
                dpred = None
        if dpred is not None:
            dpred = dpred.detach().contiguous()
            _check_dev(dpred, "dpred", (self.B, 1, self.H, self.W))
        gw: List[torch.Tensor] = []
        gb: List[torch.Tensor] = []
        if out is not None:
            gw, gb, ghw, ghb = out
            cin = self.C
            for l, (hc, k) in enumerate(zip(self.hidden, self.ksize)):
                _check_dev(gw[l], "grad weight", (4 * hc, cin + hc, k, k))
                _check_dev(gb[l], "grad bias", (4 * hc,))
                cin = hc
            _check_dev(ghw, "grad head weight", (1, self.hidden[-1], 1, 1))
            _check_dev(ghb, "grad head bias", (1,))
            if not all(t.is_contiguous() for t in [*gw, *gb, ghw, ghb]):
                raise ValueError("gradient outputs must be contiguous")
        else:
            cin = self.C
            for hc, k in zip(self.hidden, self.ksize):
                gw.append(torch.empty((4 * hc, cin + hc, k, k), dtype=torch.float32, device=self.device))
                gb.append(torch.empty((4 * hc,), dtype=torch.float32, device=self.device))
                cin = hc
            ghw = torch.empty((1, self.hidden[-1], 1, 1), dtype=torch.float32, device=self.device)
            ghb = torch.empty((1,), dtype=torch.float32, device=self.device)
        if on_ready is None:
            arr_w = (ctypes.c_void_p * self.L)(*[t.data_ptr() for t in gw])
            arr_b = (ctypes.c_void_p * self.L)(*[t.data_ptr() for t in gb])
            _lib.check(self.lib.nint_backward(self._h, _ptr(dpred), _ptr(dseq), arr_w, arr_b, _ptr(ghw), _ptr(ghb),
                                              _stream()), "nint_backward")
            return gw, gb, ghw, ghb
        # staged: the caller hears about every finished gradient bucket (the head's after BPTT, then one layer at a
        # time, top layer first) while the following wgrad kernels are still queued -- nint.h nint_backward_bptt/_wgrad
        _lib.check(self.lib.nint_backward_bptt(self._h, _ptr(dpred), _ptr(dseq), _ptr(ghw), _ptr(ghb), _stream()),
                   "nint_backward_bptt")
        on_ready(self.L)
        for l in range(self.L - 1, -1, -1):
            _lib.check(self.lib.nint_backward_wgrad(self._h, l, _ptr(gw[l]), _ptr(gb[l]), _stream()),
                       "nint_backward_wgrad")
            on_ready(l)
        return gw, gb, ghw, ghb

    # ---- measurement
    KERNEL_CLASSES = ("gate_conv_fwd", "dgrad_gate_bwd", "wgrad", "other")

    def profile(self, enable: bool):
        _lib.check(self.lib.nint_plan_profile(self._h, int(enable)), "nint_plan_profile")

    def profile_read(self):
        """{class: (device ms, launches)} since the last read (waits for the recorded events)."""
        ms = (ctypes.c_double * 4)()
        n = (ctypes.c_longlong * 4)()
        _lib.check(self.lib.nint_plan_profile_read(self._h, ms, n), "nint_plan_profile_read")
        return {k: (ms[i], n[i]) for i, k in enumerate(self.KERNEL_CLASSES)}

    def debug_raw_gates(self, x: torch.Tensor) -> torch.Tensor:
        """Gate pre-activations (no bias) of layer 0 at t=0, returned as [B,4*Hc,H,W] in the
        reference's channel order (test hook for the implicit-GEMM machinery)."""
        _check_dev(x, "x", (self.B, self.T, self.C, self.H, self.W))
        hc = self.hidden[0]
        out = torch.empty((self.B, self.H, self.W, 4 * hc), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.nint_debug_raw_gates(self._h, _ptr(x.contiguous()), _ptr(out), _stream()),
                   "nint_debug_raw_gates")
        perm = torch.tensor([self.lib.nint_gate_column(q, hc) for q in range(4 * hc)], device=self.device)
        nat = torch.empty_like(out)
        nat[..., perm] = out
        return nat.permute(0, 3, 1, 2).contiguous()
