"""torch.library registration of the ConvLSTM hot path ("a thin C-ABI torch custom-op layer").

`torch.ops.nint.convlstm_forward` / `convlstm_forward_bank` / `convlstm_backward` / `cell_forward` / `cell_backward`
are ordinary dispatcher ops with CUDA-only implementations that call libnint.so through `engine.Plan`; autograd,
fake-tensor tracing and CUDA-graph capture see them as opaque ops.  A plan (geometry, TMA descriptors, workspace with
the BPTT state) is host-side state the ops cannot carry in their arguments, so it travels as an integer handle into a
registry of live plans.  There is no CPU implementation on purpose: calling the ops with CPU tensors raises from the
dispatcher."""
import itertools
import weakref
from typing import List, Optional, Tuple

import torch
from torch import Tensor

_PLANS = weakref.WeakValueDictionary()
_ids = itertools.count(1)


def register_plan(plan) -> int:
    """Handle under which the ops find `plan`; the registry does not keep the plan alive."""
    op_id = getattr(plan, "op_id", None)
    if op_id is None:
        op_id = next(_ids)
        plan.op_id = op_id
        _PLANS[op_id] = plan
    return op_id


def _plan(plan_id: int):
    plan = _PLANS.get(plan_id)
    if plan is None:
        raise RuntimeError(f"ConvLSTM plan {plan_id} no longer exists (its module released the workspace)")
    return plan


def _set_params(plan, params):
    if len(params) != 2 * plan.L + 2:
        raise ValueError(f"expected {2 * plan.L + 2} parameter tensors, got {len(params)}")
    for l in range(plan.L):
        plan.set_weights(l, params[2 * l], params[2 * l + 1])
    plan.set_head(params[-2], params[-1])


@torch.library.custom_op("nint::convlstm_forward", mutates_args=(), device_types="cuda")
def convlstm_forward(x: Tensor, params: List[Tensor], plan_id: int) -> Tuple[Tensor, Tensor]:
    """model.py:253-274.  params = [layers.0.conv.weight, layers.0.conv.bias, ..., conv.weight, conv.bias];
    returns (pred [B,1,H,W], hs [B,T,H,W] or an empty tensor when the plan has no per-step head)."""
    plan = _plan(plan_id)
    _set_params(plan, params)
    pred, seq = plan.forward(x)
    return pred, (seq if seq is not None else pred.new_empty(0))


@convlstm_forward.register_fake
def _(x, params, plan_id):
    plan = _plan(plan_id)
    B, T, _, H, W = x.shape
    return (x.new_empty((B, 1, H, W), dtype=torch.float32),
            x.new_empty((B, T, H, W) if plan.return_sequence else (0,), dtype=torch.float32))


@torch.library.custom_op("nint::convlstm_forward_bank", mutates_args=(), device_types="cuda")
def convlstm_forward_bank(frames: Tensor, win_start: Tensor, params: List[Tensor], plan_id: int) -> Tuple[Tensor, Tensor]:
    """The same forward with the input read from an HBM-resident frame bank [N,H,W,c_pad] (dataset.py:551-637):
    sample b = frames win_start[b] .. win_start[b] + T - 1."""
    plan = _plan(plan_id)
    _set_params(plan, params)
    pred, seq = plan.forward_bank(frames, win_start)
    return pred, (seq if seq is not None else pred.new_empty(0))


@convlstm_forward_bank.register_fake
def _(frames, win_start, params, plan_id):
    plan = _plan(plan_id)
    B, H, W = win_start.shape[0], frames.shape[1], frames.shape[2]
    return (frames.new_empty((B, 1, H, W), dtype=torch.float32),
            frames.new_empty((B, plan.T, H, W) if plan.return_sequence else (0,), dtype=torch.float32))


@torch.library.custom_op("nint::convlstm_backward", mutates_args=(), device_types="cuda")
def convlstm_backward(dpred: Tensor, dseq: Optional[Tensor], plan_id: int, generation: int, need_dx: bool) -> List[Tensor]:
    """BPTT of the plan's last forward (train.py:109): gradients in the order of `params`, then dx when asked for."""
    plan = _plan(plan_id)
    if plan.generation != generation:
        raise RuntimeError("the activations this backward needs are gone: either a later forward() with the same shape "
                           "overwrote the ConvLSTM workspace, or this forward was already back-propagated once (BPTT "
                           "turns the saved gates into their gradients in place); run the forward again")
    gw, gb, ghw, ghb = plan.backward(dpred, dseq)
    grads = []
    for l in range(plan.L):
        grads += [gw[l], gb[l]]
    grads += [ghw, ghb]
    if need_dx:
        grads.append(plan.backward_input())
    return grads


@convlstm_backward.register_fake
def _(dpred, dseq, plan_id, generation, need_dx):
    plan = _plan(plan_id)
    out, cin = [], plan.C
    for hc, k in zip(plan.hidden, plan.ksize):
        out += [dpred.new_empty((4 * hc, cin + hc, k, k)), dpred.new_empty((4 * hc,))]
        cin = hc
    out += [dpred.new_empty((1, plan.hidden[-1], 1, 1)), dpred.new_empty((1,))]
    if need_dx:
        out.append(dpred.new_empty((plan.B, plan.T, plan.C, plan.H, plan.W)))
    return out


def _setup_context(ctx, inputs, output):
    ctx.plan_id = inputs[-1]
    ctx.generation = _plan(ctx.plan_id).generation


def _backward(ctx, dpred, dseq):
    plan = _plan(ctx.plan_id)
    need_dx = bool(ctx.needs_input_grad[0])
    if need_dx and not plan.input_grad:
        raise RuntimeError("gradient w.r.t. the input needs a plan made with input_grad=True (ConvLSTM.forward makes one "
                           "when x.requires_grad)")
    grads = convlstm_backward(dpred, dseq if plan.return_sequence else None, ctx.plan_id, ctx.generation, need_dx)
    dx = grads.pop() if need_dx else None
    return dx, grads, None


def _backward_bank(ctx, dpred, dseq):
    plan = _plan(ctx.plan_id)
    grads = convlstm_backward(dpred, dseq if plan.return_sequence else None, ctx.plan_id, ctx.generation, False)
    return None, None, grads, None


torch.library.register_autograd("nint::convlstm_forward", _backward, setup_context=_setup_context)
torch.library.register_autograd("nint::convlstm_forward_bank", _backward_bank, setup_context=_setup_context)


# ---- ConvLSTMCell (model.py:216-231): one fused step from an explicit state, differentiable w.r.t. x, h, c and the
# parameters like the reference's autograd module
@torch.library.custom_op("nint::cell_forward", mutates_args=(), device_types="cuda")
def cell_forward(x: Tensor, h: Tensor, c: Tensor, weight: Tensor, bias: Optional[Tensor],
                 plan_id: int) -> Tuple[Tensor, Tensor]:
    plan = _plan(plan_id)
    plan.set_weights(0, weight, bias)
    return plan.cell_forward(x, h, c)


@cell_forward.register_fake
def _(x, h, c, weight, bias, plan_id):
    return torch.empty_like(h), torch.empty_like(c)


@torch.library.custom_op("nint::cell_backward", mutates_args=(), device_types="cuda")
def cell_backward(dh: Optional[Tensor], dc: Optional[Tensor], plan_id: int, generation: int) -> List[Tensor]:
    """[dx, dh_in, dc_in, grad_weight, grad_bias] of the plan's last cell_forward."""
    plan = _plan(plan_id)
    if plan.generation != generation:
        raise RuntimeError("the activations of this ConvLSTMCell step are gone (already back-propagated once)")
    return list(plan.cell_backward(dh, dc))


@cell_backward.register_fake
def _(dh, dc, plan_id, generation):
    plan = _plan(plan_id)
    ref = dh if dh is not None else dc
    hc, k = plan.hidden[0], plan.ksize[0]
    shape = (plan.B, hc, plan.H, plan.W)
    return [ref.new_empty((plan.B, plan.C, plan.H, plan.W)), ref.new_empty(shape), ref.new_empty(shape),
            ref.new_empty((4 * hc, plan.C + hc, k, k)), ref.new_empty((4 * hc,))]


class _CellToken:
    """Dies with the autograd node of one cell step: returns the step's plan to the cell's pool."""


def _cell_setup_context(ctx, inputs, output):
    ctx.plan_id = inputs[-1]
    plan = _plan(ctx.plan_id)
    ctx.generation = plan.generation
    ctx.has_bias = inputs[4] is not None
    release = getattr(plan, "release", None)
    if release is not None:        # training plans come from ConvLSTMCell's pool (one plan per step awaiting backward)
        ctx.token = _CellToken()
        weakref.finalize(ctx.token, release)


def _cell_backward(ctx, dh, dc):
    plan = _plan(ctx.plan_id)
    if not plan.training:
        raise RuntimeError("this ConvLSTMCell step ran without autograd state (torch.no_grad or no tensor required grad)")
    if dh is None and dc is None:
        return None, None, None, None, None, None
    dx, dhi, dci, gw, gb = cell_backward(dh, dc, ctx.plan_id, ctx.generation)
    return dx, dhi, dci, gw, (gb if ctx.has_bias else None), None


torch.library.register_autograd("nint::cell_forward", _cell_backward, setup_context=_cell_setup_context)
