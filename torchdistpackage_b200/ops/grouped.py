"""Grouped (batched-over-experts) MLP on the tcgen05 GEMM: the MoE expert MLPs of one rank in six
launches (two forward, four backward) instead of six per expert.

Experts are stacked along the rows: ``x`` is ``[E * R, dim]`` (R rows per expert, the fixed-capacity
slot layout of ``moe/layer.py``), weights are ``[E, dim, hidden]`` / ``[E, hidden, dim]`` viewed as
2-D stacks.  One persistent GEMM launch walks all experts' tiles; the TMA coordinates of the
operands are shifted per expert inside the kernel (``GemmParams::grp_*``), so no data is gathered
or copied and every tile shape of the dense kernel is reused.  Weight gradients use the same
mechanism with the roles swapped (output rows = expert's weight rows, K = the expert's tokens).

STATUS: opt-in (``TDP_MOE_GROUPED=1``): the coordinate mapping is unit-tested against a per-expert
loop with an emulated kernel (tests/test_helpers.py); the kernel path has not run on hardware yet.
"""
from __future__ import annotations

import torch

from ._loader import native
from .linear import ACT_NONE, _ACT_CODE, _DACT


def _cgemm():
    return native(required=True).gemm_grouped


def grouped_supported(x: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor) -> bool:
    E, dim, hidden = w1.shape
    if x.dim() != 2 or x.shape[0] % E:
        return False
    R = x.shape[0] // E
    return (x.is_cuda and x.dtype == torch.bfloat16 and R % 128 == 0 and dim % 256 == 0
            and hidden % 256 == 0 and tuple(w2.shape) == (E, hidden, dim)
            and hasattr(native(), "gemm_grouped"))


class _GroupedMlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, act: int):
        G = _cgemm()
        E, dim, hidden = w1.shape
        R = x.shape[0] // E
        x = x.contiguous()
        z = torch.empty(E * R, hidden, dtype=x.dtype, device=x.device)
        a = torch.empty_like(z)
        y = torch.empty(E * R, dim, dtype=x.dtype, device=x.device)
        # a = act(x @ W1[e] + b1[e]);  B = W1 stacked along K ([E*dim, hidden], N contiguous)
        G(x, w1.view(E * dim, hidden), a, False, False, hidden, dim, R, b_k=dim,
          bias=None if b1 is None else b1.reshape(-1), aux_out=z, act=act)
        # y = a @ W2[e] + b2[e]
        G(a, w2.view(E * hidden, dim), y, False, False, dim, hidden, R, b_k=hidden,
          bias=None if b2 is None else b2.reshape(-1))
        ctx.save_for_backward(x, w1, w2, z, a)
        ctx.act = act
        ctx.has_bias = (b1 is not None, b2 is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        G = _cgemm()
        x, w1, w2, z, a = ctx.saved_tensors
        E, dim, hidden = w1.shape
        R = x.shape[0] // E
        dy = dy.contiguous()
        # dz = (dy @ W2[e]^T) * act'(z):  B = W2[e] read as [N = hidden, K = dim], stacked along N
        dz = torch.empty_like(z)
        G(dy, w2.view(E * hidden, dim), dz, False, True, hidden, dim, R, b_n=hidden, aux_in=z,
          act=_DACT[ctx.act])
        # dW2[e] = a_e^T @ dy_e: output rows = hidden per expert, K = the expert's R tokens
        dw2 = torch.empty_like(w2)
        G(a, dy, dw2.view(E * hidden, dim), True, False, dim, R, hidden, a_m=-hidden, a_k=R, b_k=R)
        # dx = dz @ W1[e]^T:  B = W1[e] read as [N = dim, K = hidden], stacked along N
        dx = torch.empty_like(x)
        G(dz, w1.view(E * dim, hidden), dx, False, True, dim, hidden, R, b_n=dim)
        # dW1[e] = x_e^T @ dz_e
        dw1 = torch.empty_like(w1)
        G(x, dz, dw1.view(E * dim, hidden), True, False, hidden, R, dim, a_m=-dim, a_k=R, b_k=R)
        db1 = dz.view(E, R, hidden).sum(1, dtype=torch.float32).to(dz.dtype) if ctx.has_bias[0] else None
        db2 = dy.view(E, R, dim).sum(1, dtype=torch.float32).to(dy.dtype) if ctx.has_bias[1] else None
        return dx, dw1, db1, dw2, db2, None


def grouped_mlp(x: torch.Tensor, w1: torch.Tensor, b1, w2: torch.Tensor, b2,
                act: str = "gelu_tanh") -> torch.Tensor:
    """``y[e] = act(x[e] @ w1[e] + b1[e]) @ w2[e] + b2[e]`` for the ``E`` row-groups of ``x``."""
    code = _ACT_CODE[act]
    assert code != ACT_NONE, "grouped_mlp needs an activation (gelu / gelu_tanh)"
    return _GroupedMlpFn.apply(x, w1, b1, w2, b2, code)
