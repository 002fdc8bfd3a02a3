"""Gradient clipping and loss scaling that are correct under model parallelism.

``clip_grad_norm_`` (reference: parallel/pipeline_parallel/clip_grad_parallel.py:13-77) computes the
global gradient norm across pipeline stages.  The reference sums the per-stage *norms*
(mathematically wrong for p-norms, :54-57) and ignores tensor-parallel shards and ZeRO; here:

* squared (p-th power) partial norms are summed over the ``pipe`` group, and over the ``tensor``
  group for parameters that are tensor-parallel shards (attribute ``tensor_model_parallel`` or
  ``is_tp_shard``), each replicated parameter being counted once;
* a ZeRO optimizer contributes its shard's squared norm summed over its data group;
* on GPU the per-tensor work is ONE fused sum-of-squares kernel per flat buffer / tensor and one
  fused scale kernel (csrc/fused/optim.cu), no host synchronisation (the clip coefficient stays
  on the device).

``NativeScalerPP`` is the AMP GradScaler wrapper of the reference (:100-134) with the scale /
found-inf state shared across the pipeline.
"""
from __future__ import annotations

import math
from typing import Iterable, Optional, Union

import torch
import torch.distributed as dist

from ...dist.process_topo import tpc
from ...ops._loader import native

_TensorOrTensors = Union[torch.Tensor, Iterable[torch.Tensor]]


def _is_tp_shard(p) -> bool:
    return bool(getattr(p, "tensor_model_parallel", False) or getattr(p, "is_tp_shard", False))


def _sum_pow(grads, norm_type: float, device) -> torch.Tensor:
    total = torch.zeros((), dtype=torch.float32, device=device)
    if not grads:
        return total
    C = native() if device.type == "cuda" else None
    if norm_type == 2.0 and C is not None:
        from ...ops.fused import multi_sumsq
        return multi_sumsq(list(grads))[0]          # one launch for the whole list
    for g in grads:
        total = total + g.detach().float().abs().pow(norm_type).sum()
    return total


def clip_grad_norm_(parameters: _TensorOrTensors, max_norm: float, norm_type: float = 2.0,
                    error_if_nonfinite: bool = False, foreach: Optional[bool] = None,
                    zero_optimizer=None) -> torch.Tensor:
    """Clip the *global* gradient norm (over pipe / tensor / ZeRO shards).  Returns the norm."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    params = [p for p in parameters if p.grad is not None]
    max_norm, norm_type = float(max_norm), float(norm_type)
    if not params and zero_optimizer is None:
        return torch.tensor(0.0)
    device = params[0].grad.device if params else zero_optimizer.device

    tp_on = tpc.is_mode_inited("tensor")
    pp_on = tpc.is_mode_inited("pipe")
    tp_first = (not tp_on) or tpc.get_group_rank("tensor") == 0

    if math.isinf(norm_type):
        local = torch.stack([p.grad.detach().abs().max().float() for p in params]).max() \
            if params else torch.zeros((), device=device)
        for mode in ("tensor", "pipe"):
            if tpc.is_mode_inited(mode):
                dist.all_reduce(local, op=dist.ReduceOp.MAX, group=tpc.get_group(mode))
        total_norm = local
    else:
        if zero_optimizer is not None:
            # ZeRO: every element of this rank's model lives in exactly one data-parallel shard.
            # Under tensor parallelism the sum over the tensor group must count tensor-parallel
            # *shards* on every TP rank but parameters that are *replicated* over TP (LayerNorm,
            # row-parallel bias, ...) only once: TP ranks other than the first leave them out.
            assert norm_type == 2.0, "ZeRO clipping supports the L2 norm"
            sq = zero_optimizer.local_grad_sq_norm(
                include=None if (not tp_on or tp_first) else _is_tp_shard)
            if zero_optimizer.world > 1:
                dist.all_reduce(sq, group=zero_optimizer.group)
            total = sq
            if tp_on:
                dist.all_reduce(total, group=tpc.get_group("tensor"))
        else:
            shard = [p.grad for p in params if _is_tp_shard(p)]
            repl = [p.grad for p in params if not _is_tp_shard(p)]
            total = _sum_pow(shard, norm_type, device)
            repl_pow = _sum_pow(repl, norm_type, device)
            if tp_on:
                # replicated params are identical on every TP rank: count them once
                total = total + (repl_pow if tp_first else torch.zeros_like(repl_pow))
                dist.all_reduce(total, group=tpc.get_group("tensor"))
            else:
                total = total + repl_pow
        if pp_on:
            dist.all_reduce(total, group=tpc.get_group("pipe"))
        total_norm = total.pow(1.0 / norm_type)

    if error_if_nonfinite and not torch.isfinite(total_norm):
        raise RuntimeError(f"The total norm of order {norm_type} for gradients is non-finite")
    clip_coef = torch.clamp(max_norm / (total_norm + 1e-6), max=1.0)
    if zero_optimizer is not None:
        zero_optimizer.scale_master_grads(clip_coef)
    if device.type == "cuda" and native() is not None:
        from ...ops.fused import multi_scale_
        multi_scale_([p.grad for p in params], 1.0, clip_coef)     # one launch, no host sync
    else:
        for p in params:
            p.grad.detach().mul_(clip_coef.to(p.grad.dtype))
    return total_norm


class NativeScalerPP:
    """``torch.amp.GradScaler`` front-end usable inside a pipeline: found-inf is agreed on across
    the pipe (and tensor) groups so all stages skip or take the step together."""

    state_dict_key = "amp_scaler"

    def __init__(self, enabled: bool = True, **kwargs):
        self._scaler = torch.amp.GradScaler("cuda" if torch.cuda.is_available() else "cpu",
                                            enabled=enabled, **kwargs)

    def scale(self, loss):
        return self._scaler.scale(loss)

    def _sync_found_inf(self, optimizer) -> None:
        # (a disabled GradScaler never creates its per-optimizer state)
        st = getattr(self._scaler, "_per_optimizer_states", {}).get(id(optimizer))
        if not st:
            return
        for t in st.get("found_inf_per_device", {}).values():
            for mode in ("pipe", "tensor"):
                if tpc.is_mode_inited(mode):
                    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=tpc.get_group(mode))

    def __call__(self, loss, optimizer, clip_grad=None, clip_mode: str = "norm", parameters=None,
                 create_graph=False, need_update=True, update_grad=None, backward: bool = True):
        """Positional layout of the reference (:105-114): ``(loss, optimizer, clip_grad,
        clip_mode, parameters, create_graph, need_update)``.  ``update_grad`` is timm's spelling of
        ``need_update``; ``backward=False`` when the pipeline scheduler already ran backward."""
        if clip_mode != "norm":
            raise NotImplementedError("only gradient-norm clipping is model-parallel aware")
        if update_grad is not None:
            need_update = update_grad
        if backward:
            self._scaler.scale(loss).backward(create_graph=create_graph)
        norm = None
        if need_update:
            self._scaler.unscale_(optimizer)
            self._sync_found_inf(optimizer)
            if clip_grad is not None and parameters is not None:
                norm = clip_grad_norm_(parameters, clip_grad)
            self._scaler.step(optimizer)
            self._scaler.update()
        return norm

    def state_dict(self):
        return self._scaler.state_dict()

    def load_state_dict(self, state_dict):
        self._scaler.load_state_dict(state_dict)
