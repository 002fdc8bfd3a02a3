"""Transformer MLP, serial and tensor(+sequence)-parallel.

Parity: reference ``Mlp`` / ``TpMlp`` (parallel/tensor_parallel/mlp.py:9-77): fc1 -> act -> fc2 ->
dropout; TP = column-parallel fc1 + row-parallel fc2; with sequence parallelism the input is
all-gathered first and the output reduce-scattered.

B200-first: bias + GELU live in the fc1 GEMM epilogue and the GELU derivative in the fc2 dgrad
epilogue (ops/linear.py ``mlp``); with sequence parallelism the all-gather is fused into the fc1
GEMM and the reduce-scatter into the fc2 GEMM (tp_fused.py).
"""
from __future__ import annotations

from torch import nn as nn

from ...ops import linear as _ops_linear
from .tp_utils import (ColParallelLinear, RowParallelLinear, TpLinear, get_tp_group, set_tp_group,
                       gather_from_sequence_parallel_region, reduce_from_tensor_parallel_region,
                       reduce_scatter_to_sequence_parallel_region, set_sequence_parallel_attr,
                       copy_to_tensor_parallel_region, _tp_world)
from . import tp_fused


def _act_name(act_layer) -> str:
    if act_layer is nn.GELU or isinstance(act_layer, nn.GELU):
        approx = getattr(act_layer, "approximate", "none") if isinstance(act_layer, nn.GELU) else "none"
        return "gelu_tanh" if approx == "tanh" else "gelu"
    return None


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 tp_group=None, bias=True, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        set_tp_group(tp_group)
        self.fc1 = TpLinear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self._act_name = _act_name(self.act)
        self.fc2 = TpLinear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        if self._act_name is not None:
            y = _ops_linear.mlp(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                                layout="kn", act=self._act_name)
        else:
            y = self.fc2(self.act(self.fc1(x)))
        return self.drop2(y)


class TpMlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 tp_group=None, bias=True, drop=0., sequence_parallel=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.sequence_parallel = sequence_parallel
        set_tp_group(tp_group)
        self.tp_size = _tp_world()
        self.fc1 = ColParallelLinear(in_features, hidden_features, bias=bias,
                                     input_needs_grad_reduce=not sequence_parallel)
        self.act = act_layer()
        self._act_name = _act_name(self.act)
        self.fc2 = RowParallelLinear(hidden_features, out_features, bias=bias,
                                     sequence_parallel=sequence_parallel)
        self.drop2 = nn.Dropout(drop)
        self._fused = None

    def forward(self, x):
        if self.sequence_parallel and self._act_name is not None and tp_fused.usable(x, self.tp_size):
            if self._fused is None:
                self._fused = tp_fused.FusedSpContext(get_tp_group())
            shape = x.shape
            y = tp_fused.sp_mlp(self._fused, x.reshape(-1, shape[-1]), self.fc1.linear.weight,
                                self.fc1.linear.bias, self.fc2.linear.weight, self.fc2.linear.bias,
                                self._act_name)
            return self.drop2(set_sequence_parallel_attr(y.view(*shape[:-1], y.shape[-1])))
        if self.sequence_parallel:
            x = gather_from_sequence_parallel_region(x)   # input is sequence parallel
        if self._act_name is not None:
            if not self.sequence_parallel and self.tp_size > 1:
                x = copy_to_tensor_parallel_region(x)
            y = _ops_linear.mlp(x, self.fc1.linear.weight, self.fc1.linear.bias,
                                self.fc2.linear.weight, None, layout="kn", act=self._act_name)
            y = reduce_scatter_to_sequence_parallel_region(y) if self.sequence_parallel \
                else reduce_from_tensor_parallel_region(y)
            if self.fc2.linear.bias is not None:
                y = y + self.fc2.linear.bias
                if self.sequence_parallel:
                    set_sequence_parallel_attr(y)
        else:
            y = self.fc2(self.act(self.fc1(x)))
        return self.drop2(y)
