"""Autograd front-ends of the tcgen05 GEMM (csrc/gemm): linear / fused MLP with the elementwise
work folded into the GEMM epilogues.

    linear(x, w, bias, layout, act, residual)   y = act(x @ W + b) (+ residual)
    mlp(x, w1, b1, w2, b2, ...)                 y = gelu(x @ W1 + b1) @ W2 + b2 (+ residual)

``layout='kn'`` means ``w`` is ``[in, out]`` (the reference's TpLinear convention,
tensor_parallel/tp_utils.py:162-174); ``layout='nk'`` is ``nn.Linear``'s ``[out, in]``.

Backward uses the same kernel with the other operand majors (no transposes are materialised):
dgrad = dy @ W^T, wgrad = x^T @ dy, and for the MLP the GELU derivative is applied in the epilogue
of the fc2 dgrad GEMM (``ACT_DGELU``).  bf16 CUDA tensors take the native path; everything else
(CPU, fp32) falls back to ``torch`` so the same modules run under gloo in the unit tests.
"""
from __future__ import annotations

import os
import weakref
from typing import Optional

import torch
import torch.nn.functional as F

from ._loader import native

ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_DGELU_TANH, ACT_DGELU_ERF = 0, 1, 2, 3, 4
_ACT_CODE = {None: ACT_NONE, "none": ACT_NONE, "gelu_tanh": ACT_GELU_TANH, "gelu": ACT_GELU_ERF,
             "gelu_erf": ACT_GELU_ERF}
_DACT = {ACT_GELU_TANH: ACT_DGELU_TANH, ACT_GELU_ERF: ACT_DGELU_ERF}


def _native_ok(*ts) -> bool:
    if native() is None:
        return False
    for t in ts:
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == torch.bfloat16):
            return False
    return True


def _as2d(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    return x2 if x2.is_contiguous() else x2.contiguous()


def gemm(a, b, *, trans_a=False, trans_b=False, out=None, out_dtype=None, bias=None, residual=None,
         aux_in=None, aux_out=None, act=ACT_NONE, accumulate=False, alpha=1.0, block_n=0,
         max_ctas=0, split_k=0):
    """Raw kernel entry: ``out = act(alpha * op(a) @ op(b) + bias) (+ residual)``.

    ``split_k=0`` (auto): weight-gradient shaped products (few output tiles, very long K) are
    split along K over CTAs -- partials are atomically added into an fp32 buffer and cast."""
    M = a.shape[1] if trans_a else a.shape[0]
    N = b.shape[0] if trans_b else b.shape[1]
    K = a.shape[0] if trans_a else a.shape[1]
    C = native(required=True)
    plain = (bias is None and residual is None and aux_in is None and aux_out is None
             and act == ACT_NONE and not accumulate
             and (out_dtype in (None, torch.bfloat16))
             and (out is None or (out.dtype == torch.bfloat16 and out.is_contiguous())))
    if split_k == 0 and plain:
        # weight-gradient shapes (few output tiles, K = tokens).  If 256x128 CTA-pair tiles give
        # one reasonably full wave, the 2-CTA kernel writes bf16 directly (no fp32 scratch, no
        # cast); stream-K only when the 128x256 tiles cannot fill even one wave and K is long enough
        # to amortise the fp32 atomics; everything else goes to the plain kernels (the launcher
        # picks CTA pairs for K >= 2048).
        sms = _num_sms()
        tiles = -(-M // 128) * -(-N // 256)
        pair_tiles = -(-M // 256) * -(-N // 128)
        if K >= 2048 and M > 128 and N > 128 and 0.6 * (sms // 2) <= pair_tiles <= sms // 2 \
                and _WGRAD_2CTA:
            if out is None:
                out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
            C.gemm(a, b, out, trans_a, trans_b, None, None, None, None, 0, False, float(alpha),
                   128, int(max_ctas), 1, 2)
            return out
        if tiles < sms and K >= 2048:
            split_k = 2
    if split_k > 1 and plain:
        acc = torch.zeros(M, N, dtype=torch.float32, device=a.device)
        C.gemm(a, b, acc, trans_a, trans_b, None, None, None, None, 0, False, float(alpha),
               256 if N > 128 else 128, int(max_ctas), int(split_k))
        if out is None:
            out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
        C.cast_copy(out.view(-1), acc.view(-1), 1.0)
        return out
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype or torch.bfloat16, device=a.device)
    C.gemm(a, b, out, trans_a, trans_b, bias, residual, aux_in, aux_out, int(act),
           bool(accumulate), float(alpha), int(block_n), int(max_ctas), 1)
    return out


_SMS = None
_WGRAD_2CTA = os.environ.get("TDP_WGRAD_2CTA", "1") == "1"
_FUSED_WGRAD = os.environ.get("TDP_FUSED_WGRAD", "1") == "1"


def _num_sms() -> int:
    global _SMS
    if _SMS is None:
        _SMS = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return _SMS


def colsum(x2d: torch.Tensor, out_dtype=torch.bfloat16) -> torch.Tensor:
    out = torch.empty(x2d.shape[1], dtype=out_dtype, device=x2d.device)
    if x2d.shape[1] % 8 == 0 and x2d.stride(1) == 1:
        native(required=True).colsum(x2d, out)
    else:
        out.copy_(x2d.float().sum(0))
    return out


def note_grad_stream(p: torch.Tensor) -> None:
    """Remember the stream on which ``p``'s gradient was just written into its bucket view.

    A parameter whose backward returns ``None`` (direct route) gives autograd nothing to
    synchronise: its post-accumulate hook runs with whatever stream the parameter's AccumulateGrad
    node is bound to -- the stream that was current when the node was created (e.g. a CUDA-graph
    warm-up side stream), not necessarily the stream the producing kernel was enqueued on.  The
    reducer therefore orders its comm stream behind this stream explicitly (found with a device-lag
    experiment: without it the bucket kernel could run before the weight-gradient GEMMs)."""
    if p.is_cuda:
        p._tdp_grad_stream = torch.cuda.current_stream(p.device)


def direct_grad_buffer(p: torch.Tensor):
    """``(bucket_view, overwrite)`` if parameter ``p``'s gradient can be written straight into the
    gradient bucket a reducer registered on it (see :func:`wgrad`), else ``(None, False)``."""
    buf = getattr(p, "_tdp_main_grad", None)
    if buf is None or p.grad is None or p.grad.data_ptr() != buf.data_ptr() \
            or getattr(p, "_tdp_no_fused_wgrad", False) or not _FUSED_WGRAD:
        return None, False
    return buf, bool(p._tdp_grad_fresh)


def colsum_param(bias: torch.Tensor, x2d: torch.Tensor):
    """Bias gradient ``x2d.sum(0)`` for parameter ``bias``: written straight into its gradient
    bucket view when there is one (returns ``None`` to autograd: no temporary, no ``grad += db``
    launch per bias -- 48 of those per GPT-2-small step), else returned as a tensor."""
    buf, fresh = direct_grad_buffer(bias)
    if buf is None or x2d.shape[1] % 8 != 0 or x2d.stride(1) != 1 or buf.dtype != torch.bfloat16:
        return colsum(x2d, bias.dtype)
    if fresh:
        native(required=True).colsum(x2d, buf)
        bias._tdp_grad_fresh = False
    else:
        buf.add_(colsum(x2d, buf.dtype))
    note_grad_stream(bias)
    return None


def wgrad(w: torch.Tensor, a: torch.Tensor, b: torch.Tensor):
    """Weight gradient ``a^T @ b`` for parameter ``w``.

    If ``w.grad`` is a view of a gradient bucket that a reducer (NaiveDDP) registered on the
    parameter (``w._tdp_main_grad``), the GEMM epilogue writes the bucket directly -- overwriting on
    the first micro-step after a reduction, accumulating afterwards -- and ``None`` is returned to
    autograd: no temporary, no ``grad += dw`` kernel per parameter.  Readiness is still reported by
    the reducer's post-accumulate-grad hook: autograd runs it once per parameter and backward pass
    even for an undefined gradient, after *all* uses of the parameter (tied weights) have run their
    backward.  Otherwise returns the gradient tensor for autograd to accumulate."""
    buf = getattr(w, "_tdp_main_grad", None)
    if buf is None or w.grad is None or w.grad.data_ptr() != buf.data_ptr() \
            or getattr(w, "_tdp_no_fused_wgrad", False) or not _FUSED_WGRAD:
        return gemm(a, b, trans_a=True)
    if w._tdp_grad_fresh:
        gemm(a, b, trans_a=True, out=buf)
        w._tdp_grad_fresh = False
    else:
        gemm(a, b, trans_a=True, out=buf, accumulate=True)
    note_grad_stream(w)
    return None


def _act_ref(z, act):
    if act == ACT_GELU_TANH:
        return F.gelu(z, approximate="tanh")
    if act == ACT_GELU_ERF:
        return F.gelu(z)
    return z


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, layout_nk: bool, act: int, residual):
        x2 = _as2d(x)
        N = w.shape[0] if layout_nk else w.shape[1]
        need_z = act != ACT_NONE
        z = torch.empty(x2.shape[0], N, dtype=torch.bfloat16, device=x.device) if need_z else None
        res2 = _as2d(residual) if residual is not None else None
        y = gemm(x2, w, trans_b=layout_nk, bias=bias, act=act, aux_out=z, residual=res2)
        ctx.save_for_backward(x2, w, z)
        ctx.layout_nk, ctx.act = layout_nk, act
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.bias_ref = weakref.ref(bias) if bias is not None else None
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, z = ctx.saved_tensors
        dy2 = _as2d(dy)
        if ctx.act != ACT_NONE:
            # dz = dy * act'(z): one elementwise pass folded into an identity-free path is not
            # possible here (no GEMM precedes it), so use the GEMM-free torch ops
            zf = z.float().requires_grad_(True)
            with torch.enable_grad():
                a = _act_ref(zf, ctx.act)
            dz = torch.autograd.grad(a, zf, dy2.float())[0].to(torch.bfloat16)
        else:
            dz = dy2
        dx = dw = db = dres = None
        if ctx.needs_input_grad[0]:
            # dx = dz @ W^T
            dx = gemm(dz, w, trans_b=not ctx.layout_nk).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            if ctx.layout_nk:   # dW [N, K] = dz^T @ x
                dw = wgrad(w, dz, x2)
            else:               # dW [K, N] = x^T @ dz
                dw = wgrad(w, x2, dz)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            b_ref = ctx.bias_ref() if ctx.bias_ref is not None else None
            db = colsum_param(b_ref, dz) if b_ref is not None else colsum(dz)
        if ctx.has_res and ctx.needs_input_grad[5]:
            dres = dy
        return dx, dw, db, None, None, dres


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
           layout: str = "kn", act: Optional[str] = None,
           residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    code = _ACT_CODE[act]
    K = x.shape[-1]
    N = w.shape[0] if layout == "nk" else w.shape[1]
    if _native_ok(x, w, bias, residual) and K % 8 == 0 and N % 8 == 0:
        return _LinearFn.apply(x, w, bias, layout == "nk", code, residual)
    y = torch.matmul(x, w.t() if layout == "nk" else w)
    if bias is not None:
        y = y + bias
    y = _act_ref(y, code)
    if residual is not None:
        y = y + residual
    return y


class _MlpFn(torch.autograd.Function):
    """y = act(x W1 + b1) W2 + b2 (+ residual), all elementwise work in GEMM epilogues."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, layout_nk: bool, act: int, residual):
        x2 = _as2d(x)
        H = w1.shape[0] if layout_nk else w1.shape[1]
        z = torch.empty(x2.shape[0], H, dtype=torch.bfloat16, device=x.device)
        a = gemm(x2, w1, trans_b=layout_nk, bias=b1, act=act, aux_out=z)
        res2 = _as2d(residual) if residual is not None else None
        y = gemm(a, w2, trans_b=layout_nk, bias=b2, residual=res2)
        ctx.save_for_backward(x2, w1, w2, z, a)
        ctx.layout_nk, ctx.act = layout_nk, act
        ctx.flags = (b1 is not None, b2 is not None, residual is not None)
        ctx.bias_refs = (weakref.ref(b1) if b1 is not None else None,
                         weakref.ref(b2) if b2 is not None else None)
        ctx.x_shape = x.shape
        N = w2.shape[0] if layout_nk else w2.shape[1]
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w1, w2, z, a = ctx.saved_tensors
        nk = ctx.layout_nk
        dy2 = _as2d(dy)
        has_b1, has_b2, has_res = ctx.flags
        # dz = (dy @ W2^T) * act'(z)   -- derivative applied in the epilogue
        dz = gemm(dy2, w2, trans_b=not nk, act=_DACT[ctx.act], aux_in=z)
        dw2 = wgrad(w2, dy2, a) if nk else wgrad(w2, a, dy2)
        rb1 = ctx.bias_refs[0]() if ctx.bias_refs[0] is not None else None
        rb2 = ctx.bias_refs[1]() if ctx.bias_refs[1] is not None else None
        db2 = (colsum_param(rb2, dy2) if rb2 is not None else colsum(dy2)) if has_b2 else None
        dx = gemm(dz, w1, trans_b=not nk).view(ctx.x_shape) if ctx.needs_input_grad[0] else None
        dw1 = wgrad(w1, dz, x2) if nk else wgrad(w1, x2, dz)
        db1 = (colsum_param(rb1, dz) if rb1 is not None else colsum(dz)) if has_b1 else None
        dres = dy if has_res else None
        return dx, dw1, db1, dw2, db2, None, None, dres


def mlp(x, w1, b1, w2, b2, layout: str = "kn", act: str = "gelu",
        residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    code = _ACT_CODE[act]
    if _native_ok(x, w1, b1, w2, b2, residual) and code != ACT_NONE \
            and all(d % 8 == 0 for d in (*w1.shape, *w2.shape)):
        return _MlpFn.apply(x, w1, b1, w2, b2, layout == "nk", code, residual)
    h = linear(x, w1, b1, layout, act)
    return linear(h, w2, b2, layout, None, residual)
