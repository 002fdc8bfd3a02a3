"""Tensor / sequence parallel primitives and the column / row parallel linear layers.

API parity with the reference (parallel/tensor_parallel/tp_utils.py): ``set_tp_group`` /
``get_tp_group`` (module-global group, ``None`` = world), the autograd regions
``_ReduceFromModelParallelRegion``, ``_ReduceScatterToSequenceParallelRegion``,
``_GatherFromSequenceParallelRegion``, the helpers ``gather_from_sequence_parallel_region``,
``reduce_scatter_to_sequence_parallel_region``, ``maybe_gather_from_sequence_parallel``,
``maybe_split_into_sequence_parallel``, ``set_sequence_parallel_attr``,
``is_squence_parallel_tensor`` (sic) and the layers ``TpLinear`` ([in, out] weight, ``x @ W``),
``ColParallelLinear``, ``RowParallelLinear`` with ``init_weight_from_full[_attn]``.

"Sequence" dim is dim 0 of whatever tensor is passed (reference semantics, tp_utils.py:52-108).

Differences:
* ``_CopyToModelParallelRegion`` (identity fwd / all-reduce bwd) is applied to the input of a
  column-parallel linear when sequence parallelism is off, so the *input* gradient is correct in
  plain TP (the reference leaves it a partial sum);
* ``RowParallelLinear`` adds its bias once, after the reduction (the reference adds it on every
  rank before reducing, 2.6 #12);
* bf16 CUDA tensors run the tcgen05 GEMM with fused bias epilogue (ops/linear.py); the fused
  GEMM+collective kernels are used by ``TpMlp`` / ``TpAttention`` (tp_fused.py) when sequence
  parallelism is on.  The collectives below are the generic (NCCL / gloo) path.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
from torch import nn as nn
from torch.nn.parameter import Parameter

from ...ops import linear as _ops_linear

TP_GROUP = None


def get_tp_group():
    return TP_GROUP


def set_tp_group(group):
    global TP_GROUP
    if group is not None:
        TP_GROUP = group


def reset_tp_group():
    global TP_GROUP
    TP_GROUP = None


def _tp_world() -> int:
    return dist.get_world_size(get_tp_group()) if dist.is_initialized() else 1


def _tp_rank() -> int:
    return dist.get_rank(get_tp_group()) if dist.is_initialized() else 0


def get_tensor_model_parallel_world_size() -> int:
    return _tp_world()


# ------------------------------------------------------------------------------------------
# sequence-parallel tagging
# ------------------------------------------------------------------------------------------
def set_sequence_parallel_attr(inp: torch.Tensor, value: bool = True) -> torch.Tensor:
    setattr(inp, "sequence_parallel", value)
    return inp


def is_squence_parallel_tensor(inp) -> bool:   # (sic) name kept for API compatibility
    return bool(getattr(inp, "sequence_parallel", False))


is_sequence_parallel_tensor = is_squence_parallel_tensor


# ------------------------------------------------------------------------------------------
# raw collectives along dim 0
# ------------------------------------------------------------------------------------------
def _all_reduce(x: torch.Tensor) -> torch.Tensor:
    if _tp_world() > 1:
        dist.all_reduce(x, group=get_tp_group())
    return x


def _reduce_scatter_along_first_dim(x: torch.Tensor) -> torch.Tensor:
    world = _tp_world()
    if world == 1:
        return x
    assert x.shape[0] % world == 0, "first dim must be divisible by the tensor-parallel size"
    x = x.contiguous()
    out = torch.empty((x.shape[0] // world, *x.shape[1:]), dtype=x.dtype, device=x.device)
    if x.is_cuda:
        dist.reduce_scatter_tensor(out, x, group=get_tp_group())
    else:  # gloo has no reduce_scatter
        full = x.clone()
        dist.all_reduce(full, group=get_tp_group())
        r = _tp_rank()
        out.copy_(full[r * out.shape[0]:(r + 1) * out.shape[0]])
    return out


def _gather_along_first_dim(x: torch.Tensor) -> torch.Tensor:
    world = _tp_world()
    if world == 1:
        return x
    x = x.contiguous()
    out = torch.empty((x.shape[0] * world, *x.shape[1:]), dtype=x.dtype, device=x.device)
    if x.is_cuda:
        dist.all_gather_into_tensor(out, x, group=get_tp_group())
    else:
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x, group=get_tp_group())
        out.copy_(torch.cat(parts, 0))
    return out


def _split_along_first_dim(x: torch.Tensor) -> torch.Tensor:
    world = _tp_world()
    if world == 1:
        return set_sequence_parallel_attr(x)
    assert x.shape[0] % world == 0
    k = x.shape[0] // world
    r = _tp_rank()
    return set_sequence_parallel_attr(x[r * k:(r + 1) * k].contiguous())


# ------------------------------------------------------------------------------------------
# autograd regions
# ------------------------------------------------------------------------------------------
class _CopyToModelParallelRegion(torch.autograd.Function):
    """Identity forward, all-reduce backward (input of a column-parallel layer without SP)."""

    @staticmethod
    def forward(ctx, x):
        return x

    @staticmethod
    def backward(ctx, g):
        return _all_reduce(g.contiguous().clone())


class _ReduceFromModelParallelRegion(torch.autograd.Function):
    """All-reduce forward, identity backward (output of a row-parallel layer without SP)."""

    @staticmethod
    def forward(ctx, x):
        return _all_reduce(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g


class _ReduceScatterToSequenceParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _reduce_scatter_along_first_dim(x)

    @staticmethod
    def backward(ctx, g):
        return _gather_along_first_dim(g)


class _GatherFromSequenceParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tensor_parallel_output_grad: bool = True):
        ctx.tp_out_grad = tensor_parallel_output_grad
        return _gather_along_first_dim(x)

    @staticmethod
    def backward(ctx, g):
        if ctx.tp_out_grad:
            # the gathered activation feeds tensor-parallel compute: grads are partial sums
            return _reduce_scatter_along_first_dim(g), None
        return _split_along_first_dim(g), None


def copy_to_tensor_parallel_region(x):
    return _CopyToModelParallelRegion.apply(x)


def reduce_from_tensor_parallel_region(x):
    return _ReduceFromModelParallelRegion.apply(x)


def gather_from_sequence_parallel_region(input_, tensor_parallel_output_grad: bool = True):
    out = _GatherFromSequenceParallelRegion.apply(input_, tensor_parallel_output_grad)
    return set_sequence_parallel_attr(out, False)


def reduce_scatter_to_sequence_parallel_region(input_):
    return set_sequence_parallel_attr(_ReduceScatterToSequenceParallelRegion.apply(input_))


def maybe_gather_from_sequence_parallel(inp):
    return gather_from_sequence_parallel_region(inp) if is_squence_parallel_tensor(inp) else inp


def maybe_split_into_sequence_parallel(inp):
    return inp if is_squence_parallel_tensor(inp) else _split_along_first_dim(inp)


# ------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------
class TpLinear(nn.Module):
    """``y = x @ W + b`` with ``W`` of shape ``[fin, fout]``.

    ``act`` ('gelu' / 'gelu_tanh') is fused into the GEMM epilogue on the native path."""

    def __init__(self, fin: int, fout: int, bias: bool = True, act: str = None):
        super().__init__()
        self.fin, self.fout = fin, fout
        self.weight = Parameter(torch.rand((fin, fout)))
        self.bias = Parameter(torch.zeros(fout)) if bias else None
        self.act = act

    def reset_parameters_scaled(self) -> None:
        with torch.no_grad():
            bound = 1.0 / math.sqrt(self.fin)
            self.weight.uniform_(-bound, bound)

    def forward(self, x, bias_override="default", residual=None):
        b = self.bias if isinstance(bias_override, str) else bias_override
        return _ops_linear.linear(x, self.weight, b, layout="kn", act=self.act, residual=residual)


class ColParallelLinear(nn.Module):
    """Output features split over the TP group (weight ``[fin, fout/tp]``); no communication in
    forward.  ``input_is_sequence_parallel`` is handled by the caller (gather before)."""

    def __init__(self, fin: int, fout: int, bias: bool = True, act: str = None,
                 input_needs_grad_reduce: bool = False):
        super().__init__()
        self.tp_world_size = _tp_world()
        assert fout % self.tp_world_size == 0
        self.fin = fin
        self.fout = fout // self.tp_world_size
        self.linear = TpLinear(fin, self.fout, bias, act=act)
        self.input_needs_grad_reduce = input_needs_grad_reduce
        for p in self.linear.parameters():
            p.tensor_model_parallel = True

    def forward(self, x):
        if self.input_needs_grad_reduce and self.tp_world_size > 1:
            x = copy_to_tensor_parallel_region(x)
        return self.linear(x)

    @torch.no_grad()
    def init_weight_from_full(self, fullwt: torch.Tensor, fullbias: torch.Tensor = None):
        r = _tp_rank()
        self.linear.weight.copy_(fullwt[:, r * self.fout:(r + 1) * self.fout])
        if fullbias is not None and self.linear.bias is not None:
            self.linear.bias.copy_(fullbias[r * self.fout:(r + 1) * self.fout])

    @torch.no_grad()
    def init_weight_from_full_attn(self, fullwt: torch.Tensor, fullbias: torch.Tensor = None):
        """Fused qkv weight ``[dim, 3*dim]``: slice q, k, v thirds separately so that each rank
        keeps whole heads."""
        r, ws = _tp_rank(), self.tp_world_size
        third = fullwt.shape[1] // 3
        parts = [w.split(third // ws, dim=-1)[r] for w in fullwt.split(third, dim=-1)]
        self.linear.weight.copy_(torch.cat(parts, dim=-1))
        if fullbias is not None and self.linear.bias is not None:
            bparts = [b.split(third // ws, dim=-1)[r] for b in fullbias.split(third, dim=-1)]
            self.linear.bias.copy_(torch.cat(bparts, dim=-1))


class RowParallelLinear(nn.Module):
    """Input features split over the TP group (weight ``[fin/tp, fout]``); the partial outputs
    are all-reduced, or reduce-scattered along dim 0 when ``sequence_parallel``."""

    def __init__(self, fin: int, fout: int, bias: bool = True, sequence_parallel: bool = False):
        super().__init__()
        self.tp_world_size = _tp_world()
        assert fin % self.tp_world_size == 0
        self.fin = fin // self.tp_world_size
        self.fout = fout
        self.linear = TpLinear(self.fin, fout, bias)
        self.sequence_parallel = sequence_parallel
        self.linear.weight.tensor_model_parallel = True

    def forward(self, x):
        if not self.sequence_parallel and self.tp_world_size > 1:
            # fused GEMM -> all-reduce over NVSwitch (epilogue scatter + multicast broadcast)
            from . import tp_fused
            rows = x.numel() // x.shape[-1]
            if tp_fused.usable(x, self.tp_world_size) and rows % (128 * self.tp_world_size) == 0 \
                    and self.fout % 8 == 0:
                if getattr(self, "_fused", None) is None:
                    self._fused = tp_fused.FusedSpContext(get_tp_group())
                y = tp_fused.linear_ar(self._fused, x.reshape(rows, x.shape[-1]),
                                       self.linear.weight, self.linear.bias)
                return y.view(*x.shape[:-1], self.fout)
        out = self.linear(x, bias_override=None)     # bias is added once, after the reduction
        if not self.sequence_parallel:
            out = reduce_from_tensor_parallel_region(out)
        else:
            out = reduce_scatter_to_sequence_parallel_region(out)
        if self.linear.bias is not None:
            out = out + self.linear.bias
            if self.sequence_parallel:
                set_sequence_parallel_attr(out)
        return out

    @torch.no_grad()
    def init_weight_from_full(self, fullwt: torch.Tensor, fullbias: torch.Tensor = None):
        r = _tp_rank()
        self.linear.weight.copy_(fullwt[r * self.fin:(r + 1) * self.fin])
        if fullbias is not None and self.linear.bias is not None:
            self.linear.bias.copy_(fullbias)
