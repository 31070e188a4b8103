"""Preprocessing fusion on the GPU (north-star item 4): what the reference's RNN dataset does on the host per
sample (dataset.py:520-537 stack + z-score, dataset.py:67-98 cyclic-longitude / reflect-latitude halo), as one
memory-bound kernel (`nint_fuse_inputs`).  The README's 20-level module has no shipped code: for more than one
level this follows the single-level code path and its parity is unpinned upstream."""
import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

MODES = {"reflect": 0, "reference_rnn": 1}
_DTYPES = {"bf16": (_lib.DTYPE_BF16, torch.bfloat16), "tf32": (_lib.DTYPE_TF32, torch.float32),
           "fp32": (_lib.DTYPE_TF32, torch.float32)}


def input_layout(channels: int) -> Tuple[int, int]:
    """(c_pad, ones_lane) of the model's channels-last input layout for `channels` input channels: channels padded to
    a multiple of 32 with zero lanes, the first padding lane holding 1.0 (the bias gradient rides on it; -1 when the
    channel count is already a multiple of 32).  Same rule as nint_plan_input_layout."""
    c_pad = (channels + 31) // 32 * 32
    return c_pad, (channels if c_pad > channels else -1)


class FrameBank:
    """The record resident in HBM in the model's own operand layout (the B200 form of dataset.py:551-637, where
    E33OMA90D_CRNN keeps the whole file in RAM and cuts windows with sliding_window_view): `frames` [N, Hp, Wp, c_pad]
    of bf16 (bf16 models) or tf32-rounded fp32, and optionally `targets` [N, Hy, Wy] fp32.  A training batch is then B
    window-start indices: sample b = frames[i_b : i_b + T], target = targets[i_b + T - 1] (dataset.py:600-601); the
    kernels read both through those indices, so a step uploads 4 B bytes instead of B windows."""

    def __init__(self, frames: torch.Tensor, channels: int, precision: str = "bf16", targets: Optional[torch.Tensor] = None):
        if precision not in _DTYPES:
            raise ValueError(f"precision must be one of {sorted(_DTYPES)}")
        c_pad, ones = input_layout(channels)
        if frames.dim() != 4 or frames.shape[-1] != c_pad or frames.dtype != _DTYPES[precision][1] or not frames.is_cuda:
            raise ValueError(f"frames must be a CUDA tensor [N, H, W, {c_pad}] of {_DTYPES[precision][1]}")
        self.frames, self.channels, self.c_pad, self.ones_lane, self.precision = frames.contiguous(), channels, c_pad, ones, precision
        if targets is not None:
            if targets.dim() != 3 or targets.shape[0] != frames.shape[0] or targets.dtype != torch.float32 or \
                    targets.device != frames.device:
                raise ValueError("targets must be [N, Hy, Wy] fp32 on the frames' device")
            targets = targets.contiguous()
        self.targets = targets

    def __len__(self):
        return self.frames.shape[0]

    def num_windows(self, seq_len: int) -> int:
        return len(self) - seq_len + 1

    def check_windows(self, win_start: torch.Tensor, seq_len: int):
        """Host-side bounds check when the indices are on the host (device indices are trusted: an out-of-range frame
        reads as zeros through TMA, it cannot fault)."""
        if not win_start.is_cuda and win_start.numel():
            lo, hi = int(win_start.min()), int(win_start.max())
            if lo < 0 or hi + seq_len > len(self):
                raise IndexError(f"window starts [{lo}, {hi}] with {seq_len} steps leave the bank of {len(self)} frames")

    def check_plan(self, plan):
        c_pad, ones, dt = plan.input_layout()
        if (c_pad, ones, dt) != (self.c_pad, self.ones_lane, self.frames.dtype) or plan.C != self.channels:
            raise ValueError("this frame bank was built for a different model (channels / precision)")

    @classmethod
    def from_frames(cls, x: torch.Tensor, precision: str = "bf16", targets: Optional[torch.Tensor] = None) -> "FrameBank":
        """x [N, C, H, W] fp32 or bf16 on the GPU (already normalised / padded, e.g. a dataset's X before windowing)."""
        if not x.is_cuda or x.dim() != 4 or x.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("x must be a CUDA tensor [N, C, H, W] of float32 or bfloat16")
        if x.dtype == torch.bfloat16 and precision != "bf16":
            raise TypeError("bf16 frames need precision='bf16'")
        N, C, H, W = x.shape
        c_pad, ones = input_layout(C)
        code, dt = _DTYPES[precision]
        frames = torch.empty((N, H, W, c_pad), dtype=dt, device=x.device)
        with _lib.on_device(x.device):
            _lib.check(_lib.load().nint_pack_frames(code, _lib.ptr(x.contiguous()),
                                                    _lib.X_BF16 if x.dtype == torch.bfloat16 else _lib.X_FP32, N, C, H, W,
                                                    c_pad, ones, _lib.ptr(frames), _lib.stream_ptr(x.device)),
                       "nint_pack_frames")
        return cls(frames, C, precision, targets)

    @classmethod
    def from_fields(cls, levels3d: torch.Tensor, emis2d: torch.Tensor, mean: torch.Tensor, std: torch.Tensor,
                    target_hw: Optional[Tuple[int, int]] = None, mode: str = "reflect",
                    statics: Optional[torch.Tensor] = None, precision: str = "bf16",
                    targets: Optional[torch.Tensor] = None) -> "FrameBank":
        """The preprocessing fusion (stack levels + emission, z-score, static attributes, geophysical halo) written
        straight into the bank by ONE kernel: levels3d [N, L, H, W], emis2d [N, H, W] raw fields -> frames
        [N, Hp, Wp, c_pad].  Same arithmetic as `fuse_inputs` followed by the model's input packing."""
        lead, L, H, W, S, Hp, Wp, statics = _check_fuse_args(levels3d, emis2d, mean, std, target_hw, mode, statics)
        if len(lead) != 1:
            raise ValueError("from_fields takes one leading frame dimension: levels3d [N, L, H, W]")
        C = L + 1 + S
        c_pad, ones = input_layout(C)
        if c_pad > 64:
            raise ValueError(f"{C} input channels: the fused bank kernel handles up to 64; use fuse_inputs + from_frames")
        code, dt = _DTYPES[precision]
        frames = torch.empty((lead[0], Hp, Wp, c_pad), dtype=dt, device=levels3d.device)
        with _lib.on_device(levels3d.device):
            _lib.check(_lib.load().nint_fuse_inputs_bank(_lib.ptr(levels3d.contiguous()), _lib.ptr(emis2d.contiguous()),
                                                         _lib.ptr(mean.contiguous()), _lib.ptr(std.contiguous()),
                                                         _lib.ptr(statics) if S else None, S, lead[0], L, H, W, Hp, Wp,
                                                         MODES[mode], code, c_pad, ones, _lib.ptr(frames),
                                                         _lib.stream_ptr(levels3d.device)), "nint_fuse_inputs_bank")
        return cls(frames, C, precision, targets)


def _check_fuse_args(levels3d, emis2d, mean, std, target_hw, mode, statics):
    if mode not in MODES:
        raise ValueError(f"mode must be one of {sorted(MODES)}")
    checked = [("levels3d", levels3d), ("emis2d", emis2d), ("mean", mean), ("std", std)]
    if statics is not None:
        checked.append(("statics", statics))
    for name, t in checked:
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor: the preprocessing kernel has no CPU fallback")
        if t.dtype != torch.float32:
            raise TypeError(f"{name} must be float32 (got {t.dtype})")
        if t.device != levels3d.device:
            raise RuntimeError(f"{name} is on {t.device}, levels3d on {levels3d.device}")
    lead = tuple(levels3d.shape[:-3])
    L, H, W = levels3d.shape[-3:]
    if tuple(emis2d.shape) != lead + (H, W):
        raise ValueError(f"emis2d has shape {tuple(emis2d.shape)}, expected {lead + (H, W)}")
    if mean.numel() != L + 1 or std.numel() != L + 1:
        raise ValueError(f"mean/std need {L + 1} entries")
    S = 0
    if statics is not None:
        if statics.dim() != 3 or tuple(statics.shape[1:]) != (H, W):
            raise ValueError(f"statics has shape {tuple(statics.shape)}, expected (S, {H}, {W})")
        S = statics.shape[0]
        statics = statics.contiguous()
    Hp, Wp = (H, W) if target_hw is None else (int(target_hw[0]), int(target_hw[1]))
    return lead, L, H, W, S, Hp, Wp, statics


def normalise_static_attributes(fields: torch.Tensor) -> torch.Tensor:
    """dataset.py:100-122: every static field [S, H, W] z-scored with its own spatial mean / (population) std.  Done
    once per dataset, not per sample."""
    f = fields.to(torch.float32)
    return (f - f.mean(dim=(1, 2), keepdim=True)) / f.std(dim=(1, 2), keepdim=True, unbiased=False)


def fuse_inputs(levels3d: torch.Tensor, emis2d: torch.Tensor, mean: torch.Tensor, std: torch.Tensor,
                target_hw: Optional[Tuple[int, int]] = None, mode: str = "reflect",
                statics: Optional[torch.Tensor] = None) -> torch.Tensor:
    """levels3d [..., L, H, W], emis2d [..., H, W] (same leading dims, e.g. [T] or [B, T]), mean/std [L+1],
    statics [S, H, W] (already z-scored, `normalise_static_attributes`; appended to every frame as channels L+1..,
    dataset.py:532-533) -> [..., L+1+S, Hp, Wp] fp32 on the same CUDA device.  `target_hw=None` keeps the grid size
    (no halo)."""
    lead, L, H, W, S, Hp, Wp, statics = _check_fuse_args(levels3d, emis2d, mean, std, target_hw, mode, statics)
    frames = 1
    for d in lead:
        frames *= d
    out = torch.empty(lead + (L + 1 + S, Hp, Wp), dtype=torch.float32, device=levels3d.device)
    vp = _lib.ptr
    with _lib.on_device(levels3d.device):
        _lib.check(_lib.load().nint_fuse_inputs(vp(levels3d.contiguous()), vp(emis2d.contiguous()), vp(mean.contiguous()),
                                                vp(std.contiguous()), vp(statics) if S else None, S, frames, L, H, W, Hp, Wp,
                                                MODES[mode], vp(out), _lib.stream_ptr(levels3d.device)),
                   "nint_fuse_inputs")
    return out
