"""Preprocessing fusion on the GPU (north-star item 4): what the reference's RNN dataset does on the host per
sample (dataset.py:520-537 stack + z-score, dataset.py:67-98 cyclic-longitude / reflect-latitude halo), as one
memory-bound kernel (`nint_fuse_inputs`).  The README's 20-level module has no shipped code: for more than one
level this follows the single-level code path and its parity is unpinned upstream."""
import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

MODES = {"reflect": 0, "reference_rnn": 1}


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
    frames = 1
    for d in lead:
        frames *= d
    out = torch.empty(lead + (L + 1 + S, Hp, Wp), dtype=torch.float32, device=levels3d.device)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.load().nint_fuse_inputs(vp(levels3d.contiguous()), vp(emis2d.contiguous()), vp(mean.contiguous()),
                                            vp(std.contiguous()), vp(statics) if S else None, S, frames, L, H, W, Hp, Wp,
                                            MODES[mode], vp(out), st),
               "nint_fuse_inputs")
    return out
