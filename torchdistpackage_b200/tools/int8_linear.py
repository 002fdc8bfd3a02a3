"""In-tree int8 weight-only linear (no external dependency).

The reference's module-replacement tool swaps ``nn.Linear`` for bitsandbytes / BMInf int8 layers
(tools/bnb_fc.py, tools/bminf_int8.py) to fit larger models.  Neither library is part of the B200
image, so this is the self-contained equivalent: weights are stored as int8 with one fp32 scale
per output channel (symmetric absmax quantisation), halving weight memory versus bf16; the matmul
itself runs in the activation dtype after an on-the-fly dequantisation (on a B200 the bf16 product
goes through ``ops.linear``, i.e. the tcgen05 GEMM).  Inference / frozen-weight use: the int8
weight is a buffer, not a parameter.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .module_replace import replace_all_module


class Int8WeightOnlyLinear(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.register_buffer("weight_q", torch.zeros(out_features, in_features, dtype=torch.int8))
        self.register_buffer("scale", torch.ones(out_features, dtype=torch.float32))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None

    @classmethod
    def from_linear(cls, fc: nn.Linear) -> "Int8WeightOnlyLinear":
        new = cls(fc.in_features, fc.out_features, fc.bias is not None)
        w = fc.weight.detach().float()
        scale = w.abs().amax(dim=1).clamp_min(1e-8) / 127.0
        new.weight_q = torch.round(w / scale[:, None]).clamp_(-127, 127).to(torch.int8)
        new.scale = scale
        if fc.bias is not None:
            new.bias = nn.Parameter(fc.bias.detach().clone())
        return new.to(fc.weight.device)

    def dequantized_weight(self, dtype: torch.dtype) -> torch.Tensor:
        return (self.weight_q.to(torch.float32) * self.scale[:, None]).to(dtype)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        w = self.dequantized_weight(x.dtype)
        if x.is_cuda and x.dtype == torch.bfloat16:
            from ..ops import linear as L
            return L.linear(x, w, None if self.bias is None else self.bias.to(x.dtype), layout="nk")
        return torch.nn.functional.linear(x, w, None if self.bias is None else self.bias.to(x.dtype))

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, int8 weight-only"


def replace_linear_by_int8(model: nn.Module, min_features: int = 0) -> nn.Module:
    """Swap every ``nn.Linear`` (with at least ``min_features`` inputs) for its int8 weight-only
    twin, in place; returns the model."""
    return replace_all_module(
        model, lambda m: isinstance(m, nn.Linear) and m.in_features >= min_features,
        Int8WeightOnlyLinear.from_linear)
