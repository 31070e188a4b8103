"""torch.library registration of the ConvLSTM hot path ("a thin C-ABI torch custom-op layer").

`torch.ops.nint.convlstm_forward` / `convlstm_backward` / `cell_forward` are ordinary dispatcher ops with CUDA-only
implementations that call libnint.so through `engine.Plan`; autograd, fake-tensor tracing and CUDA-graph capture
see them as opaque ops.  A plan (geometry, TMA descriptors, workspace with the BPTT state) is host-side state the
ops cannot carry in their arguments, so it travels as an integer handle into a registry of live plans.  There is no
CPU implementation on purpose: calling the ops with CPU tensors raises from the dispatcher."""
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


@torch.library.custom_op("nint::convlstm_forward", mutates_args=(), device_types="cuda")
def convlstm_forward(x: Tensor, params: List[Tensor], plan_id: int) -> Tuple[Tensor, Tensor]:
    """model.py:253-274.  params = [layers.0.conv.weight, layers.0.conv.bias, ..., conv.weight, conv.bias];
    returns (pred [B,1,H,W], hs [B,T,H,W] or an empty tensor when the plan has no per-step head)."""
    plan = _plan(plan_id)
    if len(params) != 2 * plan.L + 2:
        raise ValueError(f"expected {2 * plan.L + 2} parameter tensors, got {len(params)}")
    for l in range(plan.L):
        plan.set_weights(l, params[2 * l], params[2 * l + 1])
    plan.set_head(params[-2], params[-1])
    pred, seq = plan.forward(x)
    return pred, (seq if seq is not None else pred.new_empty(0))


@convlstm_forward.register_fake
def _(x, params, plan_id):
    plan = _plan(plan_id)
    B, T, _, H, W = x.shape
    return x.new_empty((B, 1, H, W)), x.new_empty((B, T, H, W) if plan.return_sequence else (0,))


@torch.library.custom_op("nint::convlstm_backward", mutates_args=(), device_types="cuda")
def convlstm_backward(dpred: Tensor, dseq: Optional[Tensor], plan_id: int, generation: int) -> List[Tensor]:
    """BPTT of the plan's last forward (train.py:109): gradients in the order of `params`."""
    plan = _plan(plan_id)
    if plan.generation != generation:
        raise RuntimeError("the ConvLSTM workspace was overwritten by a later forward() with the same shape "
                           "before backward() ran; run backward first (BPTT state lives in the plan workspace)")
    gw, gb, ghw, ghb = plan.backward(dpred, dseq)
    grads = []
    for l in range(plan.L):
        grads += [gw[l], gb[l]]
    return grads + [ghw, ghb]


@convlstm_backward.register_fake
def _(dpred, dseq, plan_id, generation):
    plan = _plan(plan_id)
    out, cin = [], plan.C
    for hc, k in zip(plan.hidden, plan.ksize):
        out += [dpred.new_empty((4 * hc, cin + hc, k, k)), dpred.new_empty((4 * hc,))]
        cin = hc
    return out + [dpred.new_empty((1, plan.hidden[-1], 1, 1)), dpred.new_empty((1,))]


def _setup_context(ctx, inputs, output):
    _, _, plan_id = inputs
    ctx.plan_id = plan_id
    ctx.generation = _plan(plan_id).generation


def _backward(ctx, dpred, dseq):
    if ctx.needs_input_grad[0]:
        raise NotImplementedError("gradient w.r.t. the input x is not implemented (train.py never needs it)")
    plan = _plan(ctx.plan_id)
    grads = convlstm_backward(dpred, dseq if plan.return_sequence else None, ctx.plan_id, ctx.generation)
    return None, grads, None


torch.library.register_autograd("nint::convlstm_forward", _backward, setup_context=_setup_context)


@torch.library.custom_op("nint::cell_forward", mutates_args=(), device_types="cuda")
def cell_forward(x: Tensor, h: Tensor, c: Tensor, weight: Tensor, bias: Optional[Tensor],
                 plan_id: int) -> Tuple[Tensor, Tensor]:
    """model.py:216-231: one fused cell step, (h, c) -> (h', c').  Forward-only."""
    plan = _plan(plan_id)
    plan.set_weights(0, weight, bias)
    plan.set_state(0, h, c)
    plan.forward(x.unsqueeze(1))
    return plan.get_state(0)


@cell_forward.register_fake
def _(x, h, c, weight, bias, plan_id):
    return torch.empty_like(h), torch.empty_like(c)
