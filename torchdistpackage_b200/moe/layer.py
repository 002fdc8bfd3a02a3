"""Mixture-of-experts layer with expert parallelism (NEW capability).

The reference only builds the process groups (``tpc.build_moe_groups`` -> ``moe_ep`` /
``moe_dp``, dist/process_topo.py:118-143) and the replicated-expert gradient hooks
(ddp/naive_ddp.py:233-441); the MoE layer itself is delegated to external DeepSpeed / FastMoE
forks (explore/moe/ds_fmoe_main.py).  BASELINE.json config #4 needs a real one, so here it is:

    gate (top-k softmax router, capacity factor, Switch-style balance loss)
      -> dispatch   tokens -> expert slots on the owning EP rank
      -> experts    fused MLP (tcgen05 GEMMs, bias+GELU epilogues) over each local expert's slots
      -> combine    weighted gather of expert outputs back to the token order

Dispatch / combine on B200: fixed-capacity slot layout ``[local_expert][src_rank][capacity]``
in a symmetric buffer; the source rank computes every (token, k) -> slot assignment locally
(no count exchange), and ONE kernel stores each routed row straight into the owner GPU's slot
over NVLink (``a2a_scatter_rows``); combine is the mirror image (``a2a_gather_rows``: peer loads,
gate-weight scaling and top-k accumulation fused).  Backward of dispatch is a gather, backward
of combine is a scatter (+ the gate-weight gradient dot product).  Ordering uses the in-kernel
signal-pad barrier, no NCCL.  CPU / non-symmetric groups fall back to
``dist.all_to_all_single`` on the same padded layout.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

import os

from ..ops import grouped
from ..ops import linear as L
from ..ops._loader import native
from ..ops.symm import get_symm_group


def _group_size(g) -> int:
    return dist.get_world_size(g) if dist.is_initialized() else 1


def _group_rank(g) -> int:
    return dist.get_rank(g) if dist.is_initialized() else 0


class TopKGate(nn.Module):
    """Softmax router.  Returns ``(expert_idx [T,k], weight [T,k], aux_loss)``."""

    def __init__(self, dim: int, num_experts: int, top_k: int = 2):
        super().__init__()
        self.num_experts, self.top_k = num_experts, top_k
        self.wg = nn.Parameter(torch.empty(dim, num_experts))
        nn.init.normal_(self.wg, std=0.02)

    def forward(self, x: torch.Tensor):
        logits = x.float() @ self.wg.float()
        probs = F.softmax(logits, dim=-1)
        w, idx = probs.topk(self.top_k, dim=-1)
        w = w / w.sum(-1, keepdim=True).clamp_min(1e-9)
        # Switch-transformer balance loss: E * sum_e(fraction_routed_e * mean_prob_e)
        frac = F.one_hot(idx[:, 0], self.num_experts).float().mean(0)
        aux = self.num_experts * (frac * probs.mean(0)).sum()
        return idx, w, aux


class _Plan:
    """Slot assignment of this rank's (token, k) pairs for one forward pass."""

    def __init__(self, idx: torch.Tensor, num_experts: int, ep: int, ep_rank: int, capacity: int):
        T, k = idx.shape
        e_local = num_experts // ep
        flat = idx.reshape(-1)                                     # [T*k] expert ids
        onehot = F.one_hot(flat, num_experts)                      # position within its expert
        pos = (onehot.cumsum(0) - 1).gather(1, flat[:, None]).squeeze(1)
        keep = pos < capacity
        dst_rank = (flat // e_local).to(torch.int32)
        local_e = flat % e_local
        row = (local_e * ep + ep_rank) * capacity + pos            # slot on the owner
        self.T, self.k = T, k
        self.dst_rank = dst_rank.contiguous()
        self.dst_row = torch.where(keep, row, torch.full_like(row, -1)).to(torch.int32).contiguous()
        self.keep = keep
        self.token = torch.arange(T, device=idx.device).repeat_interleave(k)
        self.capacity, self.e_local, self.ep, self.ep_rank = capacity, e_local, ep, ep_rank
        self.slots_per_rank = e_local * ep * capacity


class _A2AContext:
    """Symmetric buffers of one MoE layer (or the collective fallback)."""

    def __init__(self, group, hidden: int):
        self.group = group
        self.ep = _group_size(group)
        self.rank = _group_rank(group)
        self.hidden = hidden
        self.sym = None
        self.bufs = {}
        if self.ep > 1 and torch.cuda.is_available() and native() is not None \
                and dist.get_backend(group) == "nccl":
            sg = get_symm_group(group)
            if sg.enabled:
                self.sym = sg

    def buffer(self, name: str, rows: int):
        """Two ping-pong halves of ``rows`` slots each (zero-initialised) + the use counter."""
        key = (name, rows)
        if key not in self.bufs:
            buf = self.sym.alloc(2 * rows * self.hidden * 2)
            buf.view(0, (2 * rows, self.hidden), torch.bfloat16).zero_()
            buf.barrier(0)                  # nobody scatters into a half that is not cleared yet
            self.bufs[key] = [buf, 0]
        return self.bufs[key]


def _scatter(ctx: _A2AContext, name: str, src_rows: torch.Tensor, plan: _Plan) -> torch.Tensor:
    """rows [T*k, h] -> my slot buffer [slots_per_rank, h] filled by all EP ranks."""
    S, h = plan.slots_per_rank, src_rows.shape[1]
    if ctx.ep == 1:
        out = src_rows.new_zeros(S, h)
        ok = plan.keep
        out[plan.dst_row[ok].long()] = src_rows[ok]
        return out
    if ctx.sym is not None and src_rows.dtype == torch.bfloat16 and h % 8 == 0:
        # Ping-pong halves, ONE cross-GPU barrier per call: this call's half was cleared during
        # the previous call (before its barrier, so every peer's clear is ordered before anybody's
        # stores of this call); the other half is cleared now for the next call.  Unused slots
        # must hold zeros (the expert weight gradients sum over all slot rows).  The result is
        # copied out of symmetric memory: autograd keeps it for the expert wgrad while later
        # forwards (1F1B, micro-batching) reuse the buffer.
        entry = ctx.buffer(name, S)
        buf, use = entry[0], entry[1] & 1
        entry[1] += 1
        half = S * h * 2
        buf.view((1 - use) * half, (S, h), torch.bfloat16).zero_()
        buf.handle.a2a_scatter_rows(use * half, src_rows.contiguous(), plan.dst_rank, plan.dst_row)
        buf.barrier(entry[1] & 1)         # every rank's rows have landed (and next halves are clear)
        return buf.view(use * half, (S, h), torch.bfloat16).clone()
    # fallback: padded dense all_to_all (layout [dst_rank][local_e][src(me)][cap])
    cap, el, ep = plan.capacity, plan.e_local, ctx.ep
    send = src_rows.new_zeros(ep, el, cap, h)
    ok = plan.keep
    r = plan.dst_row[ok].long()
    le, pos = r // (ep * cap), r % cap
    send[plan.dst_rank[ok].long(), le, pos] = src_rows[ok]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(ep, -1), send.view(ep, -1), group=ctx.group)
    # recv[src][local_e][cap] -> [local_e][src][cap]
    return recv.permute(1, 0, 2, 3).reshape(S, h).contiguous()


def _gather(ctx: _A2AContext, name: str, slot_rows: torch.Tensor, plan: _Plan,
            weight: Optional[torch.Tensor], from_symm: bool) -> torch.Tensor:
    """Inverse of :func:`_scatter`: rows [T*k, h] (zeros for dropped pairs), optionally scaled."""
    h = slot_rows.shape[1]
    n = plan.T * plan.k
    if ctx.ep == 1:
        out = slot_rows.new_zeros(n, h)
        ok = plan.keep
        out[ok] = slot_rows[plan.dst_row[ok].long()]
        return out if weight is None else out * weight.reshape(-1, 1).to(out.dtype)
    if ctx.sym is not None and slot_rows.dtype == torch.bfloat16 and h % 8 == 0:
        # Ping-pong halves, ONE barrier per call: a half is rewritten two calls later, and every
        # peer has passed the barrier of the call in between only after its reads of this call.
        S = plan.slots_per_rank
        entry = ctx.buffer(name, S)
        buf, use = entry[0], entry[1] & 1
        entry[1] += 1
        half = S * h * 2
        buf.view(use * half, (S, h), torch.bfloat16).copy_(slot_rows)
        buf.barrier(entry[1] & 1)         # all ranks' slot rows are in place
        out = torch.empty(n, h, dtype=torch.bfloat16, device=slot_rows.device)
        buf.handle.a2a_gather_rows(use * half, out, plan.dst_rank, plan.dst_row,
                                   None if weight is None else weight.reshape(-1).float().contiguous(),
                                   False)
        return out
    cap, el, ep = plan.capacity, plan.e_local, ctx.ep
    send = slot_rows.view(el, ep, cap, h).permute(1, 0, 2, 3).contiguous()      # [src][le][cap]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(ep, -1), send.view(ep, -1), group=ctx.group)
    out = slot_rows.new_zeros(n, h)
    ok = plan.keep
    r = plan.dst_row[ok].long()
    out[ok] = recv[plan.dst_rank[ok].long(), r // (ep * cap), r % cap]
    return out if weight is None else out * weight.reshape(-1, 1).to(out.dtype)


class _DispatchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a2a, plan):
        ctx.a2a, ctx.plan = a2a, plan
        rows = x[plan.token]                                       # [T*k, h]
        return _scatter(a2a, "disp_fwd", rows, plan)

    @staticmethod
    def backward(ctx, d_slots):
        plan = ctx.plan
        d_rows = _gather(ctx.a2a, "disp_bwd", d_slots.contiguous(), plan, None, False)
        dx = d_rows.view(plan.T, plan.k, -1).sum(1)
        return dx.to(d_slots.dtype), None, None


class _CombineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_slots, weight, a2a, plan):
        # the un-weighted routed rows are kept for the gate-weight gradient (saves a whole
        # all-to-all in backward); the top-k weighting + sum is a local epilogue
        y_rows = _gather(a2a, "comb_fwd", y_slots.contiguous(), plan, None, False)
        ctx.a2a, ctx.plan = a2a, plan
        ctx.save_for_backward(y_rows, weight)
        w = weight.reshape(plan.T, plan.k, 1).to(y_rows.dtype)
        return (y_rows.view(plan.T, plan.k, -1) * w).sum(1)

    @staticmethod
    def backward(ctx, d_out):
        y_rows, weight = ctx.saved_tensors
        plan = ctx.plan
        d_rows = d_out[plan.token]                                 # [T*k, h]
        # d weight = <d_out[token], y_row>
        dw = (d_rows.float() * y_rows.float()).sum(-1).view(plan.T, plan.k)
        scaled = d_rows * weight.reshape(-1, 1).to(d_rows.dtype)
        d_slots = _scatter(ctx.a2a, "comb_bwd", scaled.contiguous(), plan)
        return d_slots, dw.to(weight.dtype), None, None


# all local experts in one grouped tcgen05 launch per product (validated on B200 by
# scripts/grouped_check.py / tests/test_gpu_kernels.py); TDP_MOE_GROUPED=0 restores the per-expert loop
_MOE_GROUPED = os.environ.get("TDP_MOE_GROUPED", "1") == "1"


class Experts(nn.Module):
    """``num_local`` independent MLPs with stacked weights ``[E_local, in, out]``."""

    def __init__(self, num_local: int, dim: int, hidden: int):
        super().__init__()
        self.num_local = num_local
        self.w1 = nn.Parameter(torch.empty(num_local, dim, hidden))
        self.b1 = nn.Parameter(torch.zeros(num_local, hidden))
        self.w2 = nn.Parameter(torch.empty(num_local, hidden, dim))
        self.b2 = nn.Parameter(torch.zeros(num_local, dim))
        nn.init.normal_(self.w1, std=0.02)
        nn.init.normal_(self.w2, std=0.02)

    def forward(self, slots: torch.Tensor) -> torch.Tensor:
        """slots [E_local * rows_per_expert, dim]"""
        if _MOE_GROUPED and grouped.grouped_supported(slots, self.w1, self.w2):
            # all local experts in one persistent launch per GEMM (ops/grouped.py)
            return grouped.grouped_mlp(slots, self.w1, self.b1, self.w2, self.b2, act="gelu_tanh")
        rows = slots.shape[0] // self.num_local
        outs = []
        for e in range(self.num_local):
            xe = slots[e * rows:(e + 1) * rows]
            outs.append(L.mlp(xe, self.w1[e], self.b1[e], self.w2[e], self.b2[e], layout="kn",
                              act="gelu_tanh"))
        return torch.cat(outs, 0)


class MoELayer(nn.Module):
    """Drop-in MLP replacement: ``y, aux_loss = moe(x)`` with ``x`` of shape ``[..., dim]``."""

    def __init__(self, dim: int, hidden: int, num_experts: int = 8, top_k: int = 2,
                 capacity_factor: float = 1.25, ep_group=None, expert_parallel: bool = True):
        super().__init__()
        self.ep_group = ep_group
        # expert_parallel=False keeps all experts on this rank (pure data parallel MoE)
        self.ep = _group_size(ep_group) if expert_parallel else 1
        self.ep_rank = _group_rank(ep_group) if expert_parallel else 0
        assert num_experts % self.ep == 0, "experts must divide over the expert-parallel group"
        self.num_experts, self.top_k, self.capacity_factor = num_experts, top_k, capacity_factor
        self.gate = TopKGate(dim, num_experts, top_k)
        self.experts = Experts(num_experts // self.ep, dim, hidden)
        self._a2a = None
        for p in self.experts.parameters():
            p.is_expert = True           # reduce over moe_dp, not over the full data group

    def expert_parameters(self):
        return dict(("experts." + n, p) for n, p in self.experts.named_parameters())

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        T = x2.shape[0]
        idx, w, aux = self.gate(x2)
        capacity = int(math.ceil(T * self.top_k / self.num_experts * self.capacity_factor))
        capacity = (capacity + 127) // 128 * 128      # whole GEMM tiles per (expert, source)
        plan = _Plan(idx, self.num_experts, self.ep, self.ep_rank, capacity)
        if self._a2a is None:
            self._a2a = _A2AContext(self.ep_group, shape[-1])
            if self.ep == 1:
                self._a2a.ep, self._a2a.rank, self._a2a.sym = 1, 0, None
        slots = _DispatchFn.apply(x2, self._a2a, plan)
        y_slots = self.experts(slots)
        y = _CombineFn.apply(y_slots, w.to(x2.dtype), self._a2a, plan)
        return y.view(shape), aux
