"""Host-side plan object: one `nint_plan` (include/nint.h) plus the torch-allocated workspace it
is bound to.  PyTorch is plumbing here (device memory, streams); every byte of compute happens
in libnint.so."""
import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib

PRECISIONS = {"bf16": _lib.DTYPE_BF16, "tf32": _lib.DTYPE_TF32, "fp32": _lib.DTYPE_TF32}

_ptr = _lib.ptr


class Plan:
    """Geometry + workspace for `ConvLSTM(in_channels, hidden, ksize, L)` on inputs [B,T,C,H,W].

    Every libnint call runs under `torch.cuda.device(self.device)` on torch's current stream of THAT device, and every
    tensor handed in must live on it: the library launches on the CUDA runtime's current device."""

    def __init__(self, batch: int, seq_len: int, height: int, width: int, in_channels: int,
                 hidden: Sequence[int], ksize: Sequence[int], precision: str = "bf16", training: bool = False,
                 return_sequence: bool = False, device=None, deterministic: bool = False, input_grad: bool = False):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        if len(hidden) != len(ksize):
            raise ValueError("hidden and ksize must have the same length")
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the ConvLSTM hot path has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.B, self.T, self.H, self.W, self.C = batch, seq_len, height, width, in_channels
        self.hidden, self.ksize, self.L = list(hidden), list(ksize), len(hidden)
        self.training, self.return_sequence, self.precision = bool(training), bool(return_sequence), precision
        self.deterministic, self.input_grad = bool(deterministic), bool(input_grad)
        cfg = _lib.NintConfig()
        cfg.batch, cfg.seq_len, cfg.height, cfg.width = batch, seq_len, height, width
        cfg.in_channels, cfg.num_layers = in_channels, self.L
        for i, (h, k) in enumerate(zip(hidden, ksize)):
            cfg.hidden[i], cfg.ksize[i] = h, k
        cfg.dtype, cfg.training, cfg.return_sequence = PRECISIONS[precision], int(training), int(return_sequence)
        cfg.flags = (_lib.FLAG_DETERMINISTIC if deterministic else 0) | (_lib.FLAG_INPUT_GRAD if input_grad else 0)
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.nint_plan_create(ctypes.byref(cfg), ctypes.byref(self._h)), "nint_plan_create")
        self.workspace_bytes = self.lib.nint_plan_workspace_bytes(self._h)
        with _lib.on_device(self.device):
            self.workspace = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.nint_plan_bind(self._h, _ptr(self.workspace), self.workspace_bytes, self._stream()),
                       "nint_plan_bind")
        self.generation = 0          # bumped by every forward and every backward: BPTT consumes the saved activations
        self._weight_keys = [None] * (self.L + 1)
        self._keep = None            # tensors the queued kernels still read (frame bank, window indices)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.nint_plan_destroy(h)

    def _stream(self):
        return _lib.stream_ptr(self.device)

    def _check_dev(self, t: torch.Tensor, name: str, shape=None, dtypes=(torch.float32,)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor: the ConvLSTM hot path has no CPU fallback")
        if t.device != self.device:
            raise RuntimeError(f"{name} is on {t.device} but this plan (workspace, parameters) lives on {self.device}")
        if t.dtype not in dtypes:
            raise TypeError(f"{name} must be {' or '.join(str(d) for d in dtypes)} (got {t.dtype})")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")

    def input_layout(self):
        """(c_pad, ones_lane, torch dtype) of the channels-last operand layout a frame bank for this plan must have."""
        c_pad, ones, eb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(self.lib.nint_plan_input_layout(self._h, ctypes.byref(c_pad), ctypes.byref(ones), ctypes.byref(eb)),
                   "nint_plan_input_layout")
        return c_pad.value, ones.value, (torch.bfloat16 if eb.value == 2 else torch.float32)

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
        self._check_dev(weight, "weight", (4 * hc, cin + hc, self.ksize[layer], self.ksize[layer]))
        if bias is not None:
            self._check_dev(bias, "bias", (4 * hc,))
        w = weight.detach().contiguous()
        b = None if bias is None else bias.detach().contiguous()
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_plan_set_weights(self._h, layer, _ptr(w), _ptr(b), self._stream()),
                       "nint_plan_set_weights")
        self._weight_keys[layer] = key

    def set_head(self, weight: torch.Tensor, bias: torch.Tensor, force=True):
        key = (weight.data_ptr(), weight._version, bias.data_ptr(), bias._version)
        if not force and self._weight_keys[self.L] == key:
            return
        self._check_dev(weight, "head weight", (1, self.hidden[-1], 1, 1))
        self._check_dev(bias, "head bias", (1,))
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_plan_set_head(self._h, _ptr(weight.detach().contiguous()),
                                                   _ptr(bias.detach().contiguous()), self._stream()), "nint_plan_set_head")
        self._weight_keys[self.L] = key

    # ---- state
    def reset_state(self):
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_plan_reset_state(self._h, self._stream()), "nint_plan_reset_state")

    def set_state(self, layer: int, h: torch.Tensor, c: torch.Tensor):
        shape = (self.B, self.hidden[layer], self.H, self.W)
        self._check_dev(h, "h", shape)
        self._check_dev(c, "c", shape)
        h, c = h.detach().contiguous(), c.detach().contiguous()
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_plan_set_state(self._h, layer, _ptr(h), _ptr(c), self._stream()), "nint_plan_set_state")

    def get_state(self, layer: int):
        shape = (self.B, self.hidden[layer], self.H, self.W)
        h = torch.empty(shape, dtype=torch.float32, device=self.device)
        c = torch.empty(shape, dtype=torch.float32, device=self.device)
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_plan_get_state(self._h, layer, _ptr(h), _ptr(c), self._stream()), "nint_plan_get_state")
        return h, c

    # ---- compute
    def _outputs(self):
        pred = torch.empty((self.B, 1, self.H, self.W), dtype=torch.float32, device=self.device)
        seq = (torch.empty((self.B, self.T, self.H, self.W), dtype=torch.float32, device=self.device)
               if self.return_sequence else None)
        return pred, seq

    def forward(self, x: torch.Tensor):
        """x [B,T,C,H,W] fp32, or bf16 on a bf16 plan (host-staged windows: half the PCIe bytes, identical results)."""
        self._check_dev(x, "x", (self.B, self.T, self.C, self.H, self.W), dtypes=(torch.float32, torch.bfloat16))
        if x.dtype == torch.bfloat16 and self.precision != "bf16":
            raise TypeError("bf16 inputs need precision='bf16' (a tf32 plan would lose input precision)")
        x = x.detach().contiguous()
        pred, seq = self._outputs()
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_forward_ex(self._h, _ptr(x), _lib.X_BF16 if x.dtype == torch.bfloat16 else _lib.X_FP32,
                                                _ptr(pred), _ptr(seq), self._stream()), "nint_forward")
        self.generation += 1
        return pred, seq

    def forward_bank(self, frames: torch.Tensor, win_start: torch.Tensor):
        """frames: frame bank [N,H,W,c_pad] in this plan's operand layout (preprocess.FrameBank); win_start: int32 [B]
        on the device -- sample b reads frames win_start[b] .. win_start[b]+T-1 through the TMA descriptors."""
        c_pad, _, dt = self.input_layout()
        if frames.dim() != 4 or tuple(frames.shape[1:]) != (self.H, self.W, c_pad):
            raise ValueError(f"frame bank has shape {tuple(frames.shape)}, expected (N, {self.H}, {self.W}, {c_pad})")
        self._check_dev(frames, "frame bank", dtypes=(dt,))
        self._check_dev(win_start, "win_start", (self.B,), dtypes=(torch.int32,))
        if not frames.is_contiguous() or not win_start.is_contiguous():
            raise ValueError("frame bank and win_start must be contiguous")
        pred, seq = self._outputs()
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_forward_bank(self._h, _ptr(frames), frames.shape[0], _ptr(win_start), _ptr(pred),
                                                  _ptr(seq), self._stream()), "nint_forward_bank")
        self._keep = (frames, win_start)    # the backward kernels read both again
        self.generation += 1
        return pred, seq

    def _grad_buffers(self, out):
        gw: List[torch.Tensor] = []
        gb: List[torch.Tensor] = []
        if out is not None:
            gw, gb, ghw, ghb = out
            cin = self.C
            for l, (hc, k) in enumerate(zip(self.hidden, self.ksize)):
                self._check_dev(gw[l], "grad weight", (4 * hc, cin + hc, k, k))
                self._check_dev(gb[l], "grad bias", (4 * hc,))
                cin = hc
            self._check_dev(ghw, "grad head weight", (1, self.hidden[-1], 1, 1))
            self._check_dev(ghb, "grad head bias", (1,))
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
        return gw, gb, ghw, ghb

    def backward(self, dpred: Optional[torch.Tensor], dseq: Optional[torch.Tensor] = None, out=None, on_ready=None):
        """Returns ([grad_weight_l], [grad_bias_l], grad_head_weight, grad_head_bias).  `out`: optional
        (gw list, gb list, ghw, ghb) of preallocated contiguous fp32 tensors the gradients are written into
        (e.g. views of one flat all-reduce buffer).  `on_ready(i)`: called as soon as the kernels producing a
        gradient bucket are queued on the current stream -- i = L for the head, then i = L-1 .. 0 for the layers.
        BPTT turns the saved gates into their gradients in place: one backward per forward."""
        if dseq is not None:
            dseq = dseq.detach().contiguous().clone() if dpred is not None else dseq.detach().contiguous()
            self._check_dev(dseq, "dseq", (self.B, self.T, self.H, self.W))
            if dpred is not None:   # pred is seq[:, -1]: fold its gradient in
                dseq[:, -1] += dpred.detach().reshape(self.B, self.H, self.W)
                dpred = None
        if dpred is not None:
            dpred = dpred.detach().contiguous()
            self._check_dev(dpred, "dpred", (self.B, 1, self.H, self.W))
        gw, gb, ghw, ghb = self._grad_buffers(out)
        self.generation += 1        # the saved activations are consumed: a second backward of the same forward must fail
        with _lib.on_device(self.device):
            if on_ready is None:
                arr_w = (ctypes.c_void_p * self.L)(*[t.data_ptr() for t in gw])
                arr_b = (ctypes.c_void_p * self.L)(*[t.data_ptr() for t in gb])
                _lib.check(self.lib.nint_backward(self._h, _ptr(dpred), _ptr(dseq), arr_w, arr_b, _ptr(ghw), _ptr(ghb),
                                                  self._stream()), "nint_backward")
                return gw, gb, ghw, ghb
            # staged: the caller hears about every finished gradient bucket (the head's after BPTT, then one layer at a
            # time, top layer first) while the following wgrad kernels are still queued -- nint.h nint_backward_bptt/_wgrad
            _lib.check(self.lib.nint_backward_bptt(self._h, _ptr(dpred), _ptr(dseq), _ptr(ghw), _ptr(ghb), self._stream()),
                       "nint_backward_bptt")
            on_ready(self.L)
            for l in range(self.L - 1, -1, -1):
                _lib.check(self.lib.nint_backward_wgrad(self._h, l, _ptr(gw[l]), _ptr(gb[l]), self._stream()),
                           "nint_backward_wgrad")
                on_ready(l)
        return gw, gb, ghw, ghb

    def backward_input(self) -> torch.Tensor:
        """dx [B,T,C,H,W] of the last backward (plans made with input_grad=True); call after `backward`."""
        dx = torch.empty((self.B, self.T, self.C, self.H, self.W), dtype=torch.float32, device=self.device)
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_backward_input(self._h, _ptr(dx), self._stream()), "nint_backward_input")
        return dx

    # ---- ConvLSTMCell as a differentiable unit (plans with T = 1, L = 1)
    def cell_forward(self, x, h, c):
        shape = (self.B, self.hidden[0], self.H, self.W)
        self._check_dev(x, "x", (self.B, self.C, self.H, self.W))
        self._check_dev(h, "h", shape)
        self._check_dev(c, "c", shape)
        h_out = torch.empty(shape, dtype=torch.float32, device=self.device)
        c_out = torch.empty(shape, dtype=torch.float32, device=self.device)
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_cell_forward(self._h, _ptr(x.detach().contiguous()), _ptr(h.detach().contiguous()),
                                                  _ptr(c.detach().contiguous()), _ptr(h_out), _ptr(c_out), self._stream()),
                       "nint_cell_forward")
        self.generation += 1
        return h_out, c_out

    def cell_backward(self, dh_out, dc_out, need_dx=True, need_state=True, need_params=True):
        """(dx, dh, dc, grad_weight, grad_bias) of the last cell_forward; unneeded outputs are None."""
        shape = (self.B, self.hidden[0], self.H, self.W)
        hc, k = self.hidden[0], self.ksize[0]
        for t, n in ((dh_out, "dh_out"), (dc_out, "dc_out")):
            if t is not None:
                self._check_dev(t, n, shape)
        new = lambda s: torch.empty(s, dtype=torch.float32, device=self.device)
        dx = new((self.B, self.C, self.H, self.W)) if need_dx else None
        dh = new(shape) if need_state else None
        dc = new(shape) if need_state else None
        gw = new((4 * hc, self.C + hc, k, k)) if need_params else None
        gb = new((4 * hc,)) if need_params else None
        self.generation += 1
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_cell_backward(self._h, _ptr(None if dh_out is None else dh_out.detach().contiguous()),
                                                   _ptr(None if dc_out is None else dc_out.detach().contiguous()),
                                                   _ptr(dx), _ptr(dh), _ptr(dc), _ptr(gw), _ptr(gb), self._stream()),
                       "nint_cell_backward")
        return dx, dh, dc, gw, gb

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
        self._check_dev(x, "x", (self.B, self.T, self.C, self.H, self.W))
        hc = self.hidden[0]
        if hc % 16 or (hc > 64 and hc % 64):
            raise ValueError("debug_raw_gates: hidden size must not need padding")
        out = torch.empty((self.B, self.H, self.W, 4 * hc), dtype=torch.float32, device=self.device)
        with _lib.on_device(self.device):
            _lib.check(self.lib.nint_debug_raw_gates(self._h, _ptr(x.contiguous()), _ptr(out), self._stream()),
                       "nint_debug_raw_gates")
        perm = torch.tensor([self.lib.nint_gate_column(q, hc) for q in range(4 * hc)], device=self.device)
        nat = torch.empty_like(out)
        nat[..., perm] = out
        nat = nat.permute(0, 3, 1, 2).contiguous()
        # the packed forward weights of the sigmoid gates (i, f, o) carry a factor 0.5 (the epilogue's sigmoid expects
        # halved pre-activations: exact, a power of two); undo it for the comparison with a plain convolution
        for gate in (0, 1, 3):
            nat[:, gate * hc:(gate + 1) * hc] *= 2.0
        return nat
