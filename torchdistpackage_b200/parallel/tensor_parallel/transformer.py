"""Pre-LN transformer block, serial and tensor(+sequence)-parallel, and a stack of them.

Parity: reference ``Block`` / ``ParallelBlock`` / ``Transformer``
(parallel/tensor_parallel/transformer.py:11-99): ``ln_1 -> attn -> +res -> ln_2 -> mlp -> +res``;
with sequence parallelism LN / residual run on the ``1/tp`` shard of dim 0 and attention / MLP
gather and reduce-scatter internally; ``init_from_full(blk)`` slices a serial block's weights;
``Transformer(dim, mlp_ratio, num_heads, depth, tensor_parallel, sequence_parallel)`` stacks
blocks and gathers the output at the end.

B200-first: LayerNorm and the residual add are one fused kernel (ops/fused.py ``layer_norm`` with
``residual=``), the second residual add sits in the fc2 / reduce epilogue where possible.
"""
from __future__ import annotations

import torch
from torch import nn as nn

from ...ops.fused import layer_norm
from .attn import Attention, TpAttention
from .mlp import Mlp, TpMlp
from .tp_utils import (gather_from_sequence_parallel_region, maybe_split_into_sequence_parallel,
                       set_sequence_parallel_attr)


def _ln(mod: nn.LayerNorm, x, residual=None):
    return layer_norm(x, mod.weight, mod.bias, mod.eps, residual=residual)


class Block(nn.Module):
    def __init__(self, dim, mlp_ratio=4, num_heads=8, causal=False, **not_used):
        super().__init__()
        self.ln_1 = nn.LayerNorm(dim)
        self.attn = Attention(dim, num_heads=num_heads, causal=causal)
        self.ln_2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, hidden_features=int(dim * mlp_ratio))

    def forward(self, hidden_states):
        attn_out = self.attn(_ln(self.ln_1, hidden_states))
        h, hidden_states = _ln(self.ln_2, attn_out, residual=hidden_states)   # fused add + LN
        return hidden_states + self.mlp(h)


class ParallelBlock(nn.Module):
    def __init__(self, dim, mlp_ratio=4, num_heads=8, sequence_parallel=False, causal=False):
        super().__init__()
        self.ln_1 = nn.LayerNorm(dim)
        self.attn = TpAttention(dim, num_heads=num_heads, sequence_parallel=sequence_parallel,
                                causal=causal)
        self.ln_2 = nn.LayerNorm(dim)
        self.mlp = TpMlp(dim, hidden_features=int(dim * mlp_ratio),
                         sequence_parallel=sequence_parallel)
        self.sequence_parallel = sequence_parallel
        if sequence_parallel:
            # these parameters see only 1/tp of the tokens: their grads are partial sums over the
            # tensor group (see allreduce_sequence_parallel_grads)
            for p in list(self.ln_1.parameters()) + list(self.ln_2.parameters()):
                p.sequence_parallel_grad = True
            for lin in (self.attn.proj.linear, self.mlp.fc2.linear):
                if lin.bias is not None:
                    lin.bias.sequence_parallel_grad = True

    def forward(self, hidden_states):
        if self.sequence_parallel:
            hidden_states = maybe_split_into_sequence_parallel(hidden_states)
        h = _ln(self.ln_1, hidden_states)
        if self.sequence_parallel:
            set_sequence_parallel_attr(h)
        attn_out = self.attn(h)
        h, hidden_states = _ln(self.ln_2, attn_out, residual=hidden_states)
        if self.sequence_parallel:
            set_sequence_parallel_attr(h)
        hidden_states = hidden_states + self.mlp(h)
        if self.sequence_parallel:
            set_sequence_parallel_attr(hidden_states)
        return hidden_states

    @torch.no_grad()
    def init_from_full(self, blk: Block):
        self.mlp.fc2.init_weight_from_full(blk.mlp.fc2.weight, blk.mlp.fc2.bias)
        self.mlp.fc1.init_weight_from_full(blk.mlp.fc1.weight, blk.mlp.fc1.bias)
        self.attn.qkv.init_weight_from_full_attn(blk.attn.qkv.weight, blk.attn.qkv.bias)
        self.attn.proj.init_weight_from_full(blk.attn.proj.weight, blk.attn.proj.bias)
        for mine, ref in ((self.ln_1, blk.ln_1), (self.ln_2, blk.ln_2)):
            mine.weight.copy_(ref.weight)
            mine.bias.copy_(ref.bias)


class Transformer(nn.Module):
    def __init__(self, dim, mlp_ratio=4, num_heads=8, depth=12, tensor_parallel=True,
                 sequence_parallel=True, causal=False):
        super().__init__()
        blk = ParallelBlock if tensor_parallel else Block
        self.blocks = nn.ModuleList([
            blk(dim, mlp_ratio=mlp_ratio, num_heads=num_heads, sequence_parallel=sequence_parallel,
                causal=causal) for _ in range(depth)])
        self.sequence_parallel = sequence_parallel and tensor_parallel

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        if self.sequence_parallel:
            # the gathered output feeds *replicated* compute (every TP rank evaluates the same
            # loss on it), so its gradient is identical on all ranks: backward splits it instead of
            # reduce-scattering (which would sum tp identical copies and make every gradient of the
            # stack tp times too large -- the reference does that, transformer.py:96-99)
            x = gather_from_sequence_parallel_region(x, tensor_parallel_output_grad=False)
        return x


def allreduce_sequence_parallel_grads(module: nn.Module, group=None) -> None:
    """Sum the gradients of parameters that only saw ``1/tp`` of the tokens (LayerNorm weights,
    row-parallel biases under sequence parallelism) over the tensor group.  Call after backward."""
    import torch.distributed as dist
    from .tp_utils import get_tp_group
    group = group if group is not None else get_tp_group()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in module.parameters()
             if getattr(p, "sequence_parallel_grad", False) and p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
