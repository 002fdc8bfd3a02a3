"""Attention over a packed qkv projection without layout round trips.

The attention core used for training is the library flash kernel (``scaled_dot_product_attention``:
cuDNN's sm_100 kernel on B200).  The package's own tcgen05 forward / backward kernels
(csrc/attn, ``_NativeAttnFn``) are validated on B200 (output, LSE and all three gradients against
an fp32 dense reference, deterministic; tests/test_gpu_kernels.py) but stay opt-in
(``TDP_ATTN=native``) because the library kernel is still faster at GPT-2 shapes (0.084 vs
0.068 ms forward, 0.351 vs 0.225 ms forward + backward at B=16, T=1024).
What is ours on the default path is everything around the core: q / k / v are strided *views* of the packed ``[B, T, 3*H*Dh]`` GEMM
output (no split copies), the ``[B,H,T,Dh] -> [B,T,H*Dh]`` output permute and, in backward, the
``dO`` permute and the scatter of dq / dk / dv into ONE packed ``[B, T, 3*H*Dh]`` gradient are
16-byte-vectorised row-copy kernels (csrc/fused/layout.cu) instead of the generic strided-copy
and ``cat`` kernels autograd would insert (3.5 ms of a 26 ms GPT-2-small step in the plain-torch
reference arm).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from ._loader import native


class _PackedAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, n_head: int, causal: bool, scale: Optional[float]):
        C = native()
        B, T, D3 = qkv.shape
        D = D3 // 3
        dh = D // n_head
        qkv5 = qkv.view(B, T, 3, n_head, dh)
        with torch.enable_grad():
            q, k, v = (qkv5[:, :, i].transpose(1, 2).detach().requires_grad_(True) for i in range(3))
            o = F.scaled_dot_product_attention(q, k, v, is_causal=causal, scale=scale)
        ot = o.transpose(1, 2)
        if ot.is_contiguous():
            # the library laid O out like Q (physically [B, T, H, Dh]): already what we return
            out = ot.detach()
        else:
            out = torch.empty(B, T, n_head, dh, dtype=qkv.dtype, device=qkv.device)
            C.permute_rows_copy(out.transpose(1, 2), o)
        ctx.graph = (q, k, v, o)
        ctx.dims = (B, T, n_head, dh)
        return out.view(B, T, D)

    @staticmethod
    def backward(ctx, dout):
        C = native()
        B, T, H, dh = ctx.dims
        q, k, v, o = ctx.graph
        ctx.graph = None
        dview = dout.reshape(B, T, H, dh).transpose(1, 2)
        if dview.stride() == o.stride():
            do = dview                      # dO already has O's physical layout: no permute
        else:
            do = torch.empty(B, H, T, dh, dtype=dout.dtype, device=dout.device)
            C.permute_rows_copy(do, dview)
        dq, dk, dv = torch.autograd.grad(o, (q, k, v), do)
        dqkv = torch.empty(B, T, 3, H, dh, dtype=dout.dtype, device=dout.device)
        gs = [g if g.stride(3) == 1 else g.contiguous() for g in (dq, dk, dv)]
        dsts = [dqkv[:, :, i].transpose(1, 2) for i in range(3)]
        # one launch packs all three (falls back to one launch each if their strides differ)
        if not (hasattr(C, "permute_rows_copy3") and C.permute_rows_copy3(dsts, gs)):
            for d, g in zip(dsts, gs):
                C.permute_rows_copy(d, g)
        return dqkv.view(B, T, 3 * H * dh), None, None, None


def native_attention_forward(qkv: torch.Tensor, n_head: int, causal: bool = True,
                             scale: Optional[float] = None, return_lse: bool = False):
    """Forward-only attention on the package's own tcgen05 kernel (csrc/attn/attn_fwd_sm100.cu):
    q, k, v are read in place from the packed projection, the output is written in ``[B, T, H*Dh]``
    layout, no layout copies at all.  head_dim 64, ``T % 128 == 0``.  (``_NativeAttnFn`` adds the
    backward kernel; both are selected with ``TDP_ATTN=native``.)"""
    C = native(required=True)
    B, T, D3 = qkv.shape
    D = D3 // 3
    dh = D // n_head
    qkv5 = qkv.detach().view(B, T, 3, n_head, dh)
    out = torch.empty(B, T, n_head, dh, dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty(B, n_head, T, dtype=torch.float32, device=qkv.device) if return_lse else None
    C.attn_fwd(qkv5[:, :, 0], qkv5[:, :, 1], qkv5[:, :, 2], out, lse, bool(causal),
               float(scale if scale is not None else dh ** -0.5))
    out = out.view(B, T, D)
    return (out, lse) if return_lse else out


class _NativeAttnFn(torch.autograd.Function):
    """Forward and backward on the package's tcgen05 attention kernels.  q, k, v are read in place
    from the packed projection; dq, dk and dv are written by TMA straight into their column
    windows of the packed gradient (dQ and dK/dV each have their own output-stationary kernel, so
    nothing is accumulated through global memory)."""

    @staticmethod
    def forward(ctx, qkv, n_head: int, causal: bool, scale: float):
        out, lse = native_attention_forward(qkv, n_head, causal, scale, return_lse=True)
        ctx.save_for_backward(qkv, out, lse)
        ctx.cfg = (n_head, causal, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        C = native(required=True)
        qkv, out, lse = ctx.saved_tensors
        H, causal, scale = ctx.cfg
        B, T, D3 = qkv.shape
        dh = D3 // 3 // H
        dout = dout.contiguous()
        do4, o4 = dout.view(B, T, H, dh), out.view(B, T, H, dh)
        qkv5 = qkv.view(B, T, 3, H, dh)
        dqkv = torch.empty_like(qkv)
        dqkv5 = dqkv.view(B, T, 3, H, dh)
        delta = torch.empty(B, H, T, dtype=torch.float32, device=qkv.device)   # scratch
        # two launches: (dQ + delta) then (dK, dV); all three land in their windows of dqkv
        C.attn_bwd(qkv5[:, :, 0], qkv5[:, :, 1], qkv5[:, :, 2], o4, do4, lse, delta,
                   dqkv5[:, :, 0], dqkv5[:, :, 1], dqkv5[:, :, 2], bool(causal), float(scale))
        return dqkv, None, None, None


def _native_attention_enabled(qkv: torch.Tensor, n_head: int) -> bool:
    import os
    if os.environ.get("TDP_ATTN", "library") != "native":
        return False
    B, T, D3 = qkv.shape
    dh = D3 // 3 // n_head
    return dh == 64 and T % 128 == 0 and qkv.dtype == torch.bfloat16 and hasattr(native(), "attn_bwd")


def packed_attention(qkv: torch.Tensor, n_head: int, causal: bool = True,
                     scale: Optional[float] = None) -> torch.Tensor:
    """``qkv`` ``[B, T, 3*H*Dh]`` (q | k | v along the last dim) -> ``[B, T, H*Dh]``."""
    B, T, D3 = qkv.shape
    D = D3 // 3
    dh = D // n_head
    if native() is not None and qkv.is_cuda and qkv.is_contiguous() \
            and _native_attention_enabled(qkv, n_head):
        sc = float(scale if scale is not None else dh ** -0.5)
        if torch.is_grad_enabled() and qkv.requires_grad:
            return _NativeAttnFn.apply(qkv, n_head, causal, sc)
        return native_attention_forward(qkv, n_head, causal, sc)
    if native() is not None and qkv.is_cuda and qkv.dtype in (torch.bfloat16, torch.float16) \
            and (dh * qkv.element_size()) % 16 == 0 and qkv.is_contiguous():
        return _PackedAttnFn.apply(qkv, n_head, causal, scale)
    q, k, v = qkv.view(B, T, 3, n_head, dh).permute(2, 0, 3, 1, 4)
    o = F.scaled_dot_product_attention(q, k, v, is_causal=causal, scale=scale)
    return o.transpose(1, 2).reshape(B, T, D)
