"""Multi-head self attention, serial and tensor(+sequence)-parallel.

Parity: reference ``Attention`` / ``TpAttention`` (parallel/tensor_parallel/attn.py:18-98):
fused qkv projection (``TpLinear(dim, 3*dim)`` / column parallel with head-aligned slicing),
``heads/tp`` local heads, output projection (row parallel), no mask by default (ViT style).

B200-first: the score matrix is never materialised -- the core runs through
``scaled_dot_product_attention`` (flash kernel, library call) in ``[B, h, N, d]`` layout; the qkv
and proj GEMMs are the tcgen05 kernel; with sequence parallelism the all-gather -> qkv GEMM and
proj GEMM -> reduce-scatter pairs are the fused GEMM+collective kernels (tp_fused.py).
``causal=True`` is an extension used by the GPT models.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn as nn

from .tp_utils import (ColParallelLinear, RowParallelLinear, TpLinear, get_tp_group, set_tp_group,
                       gather_from_sequence_parallel_region, set_sequence_parallel_attr, _tp_world)
from . import tp_fused


def _split_heads(t: torch.Tensor, num_heads: int, head_dim: int) -> torch.Tensor:
    """[B, N, h*d] -> [B, h, N, d]"""
    return t.view(*t.shape[:-1], num_heads, head_dim).permute(0, 2, 1, 3)


def attention_core(q, k, v, scale: float, dropout_p: float = 0.0, causal: bool = False,
                   training: bool = True):
    """softmax(q k^T * scale) v for [B, h, N, d] tensors without materialising [N, N]."""
    return F.scaled_dot_product_attention(q, k, v, dropout_p=dropout_p if training else 0.0,
                                          is_causal=causal, scale=scale)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0., causal=False):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.causal = causal
        self.qkv = TpLinear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = TpLinear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, D = x.shape
        q, k, v = self.qkv(x).chunk(3, dim=-1)
        q = _split_heads(q, self.num_heads, self.head_dim)
        k = _split_heads(k, self.num_heads, self.head_dim)
        v = _split_heads(v, self.num_heads, self.head_dim)
        o = attention_core(q, k, v, self.scale, self.attn_drop.p, self.causal, self.training)
        o = o.transpose(1, 2).reshape(B, N, D)
        return self.proj_drop(self.proj(o))


class TpAttention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.,
                 tp_group=None, sequence_parallel=False, causal=False):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        set_tp_group(tp_group)
        self.tp_size = _tp_world()
        assert num_heads % self.tp_size == 0, "heads must divide over the tensor-parallel group"
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.head_num_per_partition = num_heads // self.tp_size
        self.causal = causal
        self.sequence_parallel = sequence_parallel
        self.qkv = ColParallelLinear(dim, dim * 3, bias=qkv_bias,
                                     input_needs_grad_reduce=not sequence_parallel)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = RowParallelLinear(dim, dim, sequence_parallel=sequence_parallel)
        self.proj_drop = nn.Dropout(proj_drop)
        self._fused = None

    def _core(self, qkv_out, B, N):
        hp, d = self.head_num_per_partition, self.head_dim
        q, k, v = qkv_out.view(B, N, -1).chunk(3, dim=-1)
        q, k, v = (_split_heads(t, hp, d) for t in (q, k, v))
        o = attention_core(q, k, v, self.scale, self.attn_drop.p, self.causal, self.training)
        return o.transpose(1, 2).reshape(B, N, hp * d)

    def forward(self, x):
        if self.sequence_parallel and tp_fused.usable(x, self.tp_size):
            # fused all-gather->qkv GEMM and proj GEMM->reduce-scatter (+bias) over NVSwitch
            if self._fused is None:
                self._fused = tp_fused.FusedSpContext(get_tp_group())
            Bs, N, D = x.shape
            B = Bs * self.tp_size
            qkv = tp_fused.ag_linear(self._fused, x.reshape(Bs * N, D), self.qkv.linear.weight,
                                     self.qkv.linear.bias, slot="attn_in")
            o = self._core(qkv, B, N)
            y = tp_fused.linear_rs(self._fused, o.reshape(B * N, -1), self.proj.linear.weight,
                                   self.proj.linear.bias, slot="attn_out")
            return self.proj_drop(set_sequence_parallel_attr(y.view(Bs, N, D)))
        if self.sequence_parallel:
            x = gather_from_sequence_parallel_region(x)   # input is sequence parallel
        B, N, D = x.shape
        o = self._core(self.qkv(x), B, N)
        return self.proj_drop(self.proj(o))
