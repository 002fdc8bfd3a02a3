"""Autograd / functional front-ends of the HBM-bound fused kernels (csrc/fused/*.cu) with
PyTorch fallbacks for CPU tensors.

    layer_norm(x, weight, bias, eps, residual=None) -> y  or (y, x + residual)
    cross_entropy(logits, target, ignore_index)     -> mean loss  (fwd+bwd in ONE pass: the
                                                       logits buffer is overwritten with dlogits)
    FusedAdamW                                       flat multi-tensor AdamW driven by one launch
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F

from ._loader import native


def dist_rank(group=None) -> int:
    import torch.distributed as dist
    return dist.get_rank(group) if dist.is_initialized() else 0


def _fused_ok(*ts) -> bool:
    if native() is None:
        return False
    return all(t is None or (t.is_cuda and t.dtype == torch.bfloat16) for t in ts)



def _ln_param_grads(weight, bias_p, has_bias, cols, dev, wdtype):
    """Output tensors for LayerNorm's dgamma / dbeta: the parameters' gradient-bucket views when a
    reducer registered them and they are fresh (returns flags telling what autograd should get)."""
    from .linear import direct_grad_buffer
    gbuf, gfresh = direct_grad_buffer(weight) if weight is not None else (None, False)
    bbuf, bfresh = direct_grad_buffer(bias_p) if (bias_p is not None and has_bias) else (None, False)
    direct = gbuf is not None and gfresh and gbuf.dtype == wdtype and \
        (not has_bias or (bbuf is not None and bfresh and bbuf.dtype == wdtype))
    if direct:
        from .linear import note_grad_stream
        dgamma = gbuf
        dbeta = bbuf if has_bias else torch.empty(cols, dtype=wdtype, device=dev)
        weight._tdp_grad_fresh = False
        note_grad_stream(weight)
        if has_bias:
            bias_p._tdp_grad_fresh = False
            note_grad_stream(bias_p)
        return dgamma, dbeta, True
    return (torch.empty(cols, dtype=wdtype, device=dev), torch.empty(cols, dtype=wdtype, device=dev),
            False)


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps: float, residual):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous()
        r2 = residual.reshape(-1, shape[-1]).contiguous() if residual is not None else None
        rows, cols = x2.shape
        y = torch.empty_like(x2)
        s = torch.empty_like(x2) if residual is not None else None
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        native().layernorm_fwd(x2, r2, weight, bias, y, s, mean, rstd, float(eps))
        ctx.save_for_backward(s if residual is not None else x2, weight, mean, rstd)
        ctx.has_res, ctx.has_bias = residual is not None, bias is not None
        import weakref
        ctx.w_ref = weakref.ref(weight)
        ctx.b_ref = weakref.ref(bias) if bias is not None else None
        ctx.shape = shape
        if residual is not None:
            return y.view(shape), s.view(shape)
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy, ds=None):
        xin, weight, mean, rstd = ctx.saved_tensors
        rows, cols = xin.shape
        dy2 = dy.reshape(rows, cols).contiguous()
        ds2 = ds.reshape(rows, cols).contiguous() if (ctx.has_res and ds is not None) else None
        dx = torch.empty_like(xin)
        dgamma, dbeta, direct = _ln_param_grads(ctx.w_ref(), ctx.b_ref() if ctx.b_ref else None,
                                                ctx.has_bias, cols, xin.device, weight.dtype)
        native().layernorm_bwd(dy2, xin, weight, mean, rstd, dx, ds2, dgamma, dbeta)
        dxv = dx.view(ctx.shape)
        if direct:
            return dxv, None, None, None, (dxv if ctx.has_res else None)
        return dxv, dgamma, (dbeta if ctx.has_bias else None), None, (dxv if ctx.has_res else None)


class _LayerNormForkFn(torch.autograd.Function):
    """``(LN(x), x)``: the pre-LN residual fork of a transformer block as ONE autograd node.

    ``x`` feeds both the LayerNorm and the residual add further down; autograd would sum the two
    gradient contributions with a separate element-wise kernel (24 of them per GPT-2-small step).
    Here the second output is ``x`` itself (an alias), and backward hands both incoming gradients
    to the LayerNorm backward kernel, which adds the skip gradient while it writes ``dx``."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps: float):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        rows, cols = x2.shape
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        native().layernorm_fwd(x2, None, weight, bias, y, None, mean, rstd, float(eps))
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.has_bias = bias is not None
        import weakref
        ctx.w_ref = weakref.ref(weight)
        ctx.b_ref = weakref.ref(bias) if bias is not None else None
        ctx.shape = shape
        return y.view(shape), x.view(shape)

    @staticmethod
    def backward(ctx, dy, dskip):
        xin, weight, mean, rstd = ctx.saved_tensors
        rows, cols = xin.shape
        if dy is None:                       # only the skip path was used
            return dskip, None, None, None
        dy2 = dy.reshape(rows, cols).contiguous()
        ds2 = dskip.reshape(rows, cols).contiguous() if dskip is not None else None
        dx = torch.empty_like(xin)
        dgamma, dbeta, direct = _ln_param_grads(ctx.w_ref(), ctx.b_ref() if ctx.b_ref else None,
                                                ctx.has_bias, cols, xin.device, weight.dtype)
        native().layernorm_bwd(dy2, xin, weight, mean, rstd, dx, ds2, dgamma, dbeta)
        if direct:
            return dx.view(ctx.shape), None, None, None
        return dx.view(ctx.shape), dgamma, (dbeta if ctx.has_bias else None), None


def layer_norm_fork(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                    eps: float = 1e-5):
    """``h, skip = layer_norm_fork(x, w, b)``: ``h = LN(x)`` and ``skip`` is ``x`` for the residual
    connection; use ``skip`` (not ``x``) downstream so that the two gradients meet inside the
    LayerNorm backward kernel instead of in a separate add."""
    cols = x.shape[-1]
    if _fused_ok(x, weight, bias, None) and cols % 8 == 0 and cols <= 8192 and x.is_contiguous():
        return _LayerNormForkFn.apply(x, weight, bias, eps)
    return F.layer_norm(x, (cols,), weight, bias, eps), x


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
               eps: float = 1e-5, residual: Optional[torch.Tensor] = None):
    """LayerNorm over the last dim.  With ``residual`` returns ``(LN(x + residual), x + residual)``
    (the sum is produced by the same kernel: one read of each operand, two writes)."""
    cols = x.shape[-1]
    if _fused_ok(x, weight, bias, residual) and cols % 8 == 0 and cols <= 8192:
        return _LayerNormFn.apply(x, weight, bias, eps, residual)
    if residual is not None:
        s = x + residual
        return F.layer_norm(s, (cols,), weight, bias, eps), s
    return F.layer_norm(x, (cols,), weight, bias, eps)


class _CrossEntropyFn(torch.autograd.Function):
    """Mean cross entropy whose backward was already computed in forward (dlogits overwrite the
    logits).  ``logits`` must therefore not be used afterwards -- it is the LM-head output."""

    @staticmethod
    def forward(ctx, logits2d, target, ignore_index: int):
        rows = logits2d.shape[0]
        n_valid = (target != ignore_index).sum().clamp_min(1)
        loss_rows = torch.empty(rows, dtype=torch.float32, device=logits2d.device)
        # grad wrt the *mean*: scale by 1/n_valid; done on device without a sync by a second pass
        native().cross_entropy_fwd_bwd(logits2d, target, loss_rows, 1.0, int(ignore_index))
        inv = (1.0 / n_valid.float()).reshape(1)
        ctx.save_for_backward(logits2d, inv)
        return loss_rows.sum() * inv[0]

    @staticmethod
    def backward(ctx, dloss):
        dlogits, inv = ctx.saved_tensors
        coef = (dloss.float().reshape(1) * inv)
        native().scale_(dlogits.view(-1) if dlogits.is_contiguous() else dlogits, 1.0, coef)
        return dlogits, None, None


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = -100):
    """Mean softmax cross entropy over ``[..., vocab]`` logits."""
    vocab = logits.shape[-1]
    l2 = logits.reshape(-1, vocab)
    t = target.reshape(-1)
    if _fused_ok(l2) and l2.is_contiguous() and logits.requires_grad and vocab % 8 == 0:
        return _CrossEntropyFn.apply(l2, t.contiguous(), ignore_index)
    return F.cross_entropy(l2.float(), t, ignore_index=ignore_index)


class _LmHeadLossFn(torch.autograd.Function):
    """LM head + softmax cross entropy as ONE terminal op: logits = h @ W^T (tcgen05 GEMM), fused
    CE kernel overwrites them with d(logits) (already scaled by 1/num_valid), and the two
    backward GEMMs (d_h = dlogits @ W, d_W = dlogits^T @ h) run right away while the logits are
    hot -- the [tokens, vocab] tensor never outlives forward and is touched 3 times instead of 6.
    Backward only rescales the two small gradients by the incoming d(loss)."""

    @staticmethod
    def forward(ctx, hidden, weight, target, ignore_index: int, num_valid: int):
        from . import linear as L
        C = native()
        D = hidden.shape[-1]
        h2 = hidden.reshape(-1, D)
        h2 = h2 if h2.is_contiguous() else h2.contiguous()
        logits = L.gemm(h2, weight, trans_b=True)                       # [T, V]
        loss_rows = torch.empty(h2.shape[0], dtype=torch.float32, device=h2.device)
        C.cross_entropy_fwd_bwd(logits, target.reshape(-1).contiguous(), loss_rows,
                                1.0 / float(num_valid), int(ignore_index))
        d_h = L.gemm(logits, weight)                                    # [T, D]
        d_w = L.gemm(logits, h2, trans_a=True)                          # [V, D]
        ctx.save_for_backward(d_h, d_w)
        ctx.h_shape = hidden.shape
        import weakref
        ctx.weight_ref = weakref.ref(weight)
        return loss_rows.sum() / float(num_valid)

    @staticmethod
    def backward(ctx, dloss):
        d_h, d_w = ctx.saved_tensors
        C = native()
        s = dloss.reshape(1).float()
        C.scale_(d_h.view(-1), 1.0, s)
        C.scale_(d_w.view(-1), 1.0, s)
        # tied embedding: the embedding backward (last op of backward) adds its rows straight into
        # this dense gradient (`tied_embedding`) instead of building a second [vocab, d] tensor
        w = ctx.weight_ref()
        if w is not None:
            w._tdp_pending_dw = d_w
        return d_h.view(ctx.h_shape), d_w, None, None, None


class _TiedEmbeddingFn(torch.autograd.Function):
    """Embedding lookup whose weight is tied to the LM head (GPT-2).  Backward scatters d(out)
    rows into the LM head's dense weight gradient, which ``lm_head_loss`` left on the parameter:
    one row-scatter kernel with bf16x2 atomics replaces autograd's zero-filled [vocab, d] tensor,
    the sort-based embedding backward and the dense add of the two contributions."""

    @staticmethod
    def forward(ctx, idx, weight):
        ctx.save_for_backward(idx)
        import weakref
        ctx.weight_ref = weakref.ref(weight)
        ctx.wshape = weight.shape
        return F.embedding(idx, weight)

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        w = ctx.weight_ref()
        pend = getattr(w, "_tdp_pending_dw", None) if w is not None else None
        d2 = dout.reshape(-1, dout.shape[-1])
        if pend is not None:
            del w._tdp_pending_dw
        C = native()
        if pend is not None and C is not None and pend.is_cuda and pend.dtype == torch.bfloat16 \
                and pend.shape == ctx.wshape and d2.dtype == torch.bfloat16 and pend.is_contiguous() \
                and (pend.shape[1] * 2) % 16 == 0:
            C.rows_scatter_add(pend, idx.reshape(-1).contiguous(), d2.contiguous())
            return None, None            # already part of the LM head's gradient tensor
        dw = torch.zeros(ctx.wshape, dtype=dout.dtype, device=dout.device)
        dw.index_add_(0, idx.reshape(-1), d2)
        return None, dw


def tied_embedding(idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``F.embedding(idx, weight)`` for a weight that is also the LM head of ``lm_head_loss``."""
    return _TiedEmbeddingFn.apply(idx, weight)


def lm_head_loss(hidden: torch.Tensor, weight: torch.Tensor, target: torch.Tensor,
                 ignore_index: Optional[int] = None) -> torch.Tensor:
    """Mean next-token cross entropy of ``hidden @ weight^T`` (``weight`` = tied embedding
    ``[vocab, dim]``).  ``ignore_index=None`` means every target counts (no host sync); with an
    ignore index the number of valid targets is read back once."""
    V, D = weight.shape
    if _fused_ok(hidden, weight) and V % 8 == 0 and D % 8 == 0 and hidden.requires_grad:
        if ignore_index is None:
            n_valid, ign = target.numel(), -100
            # (-100 never occurs in token ids, so nothing is ignored)
        else:
            n_valid, ign = int((target != ignore_index).sum().item()), int(ignore_index)
        return _LmHeadLossFn.apply(hidden, weight, target, ign, max(n_valid, 1))
    logits = torch.matmul(hidden, weight.t())
    return F.cross_entropy(logits.reshape(-1, V).float(), target.reshape(-1),
                           ignore_index=-100 if ignore_index is None else ignore_index)


class FusedAdamW(torch.optim.Optimizer):
    """AdamW over *flat* buffers: every parameter (and its grad) of a group is a view into one
    contiguous buffer, so ``step()`` is a single kernel launch per group
    (csrc/fused/optim.cu ``adamw_kernel``) with fp32 master weights + moments and bf16 write-out.

    Use :func:`flatten_module_params` (or NaiveDDP's bucket views for the grads) to obtain flat
    storage; parameters that are not flat-backed fall back to per-tensor launches.
    """

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 master_weights: bool = True, adamw_mode: bool = True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.master_weights = master_weights
        self.adamw_mode = adamw_mode
        self._flat = {}     # group index -> (flat_param, flat_grad or None)

    def attach_flat(self, group_index: int, flat_param: torch.Tensor,
                    flat_grad: Optional[torch.Tensor] = None) -> None:
        """Declare that all params of ``group_index`` are views of ``flat_param`` (and their grads
        views of ``flat_grad`` at the same offsets)."""
        self._flat[group_index] = (flat_param, flat_grad)

    def _state_for(self, key, like: torch.Tensor):
        st = self.state[key]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros(like.numel(), dtype=torch.float32, device=like.device)
            st["exp_avg_sq"] = torch.zeros(like.numel(), dtype=torch.float32, device=like.device)
            if self.master_weights and like.dtype != torch.float32:
                st["master"] = like.detach().reshape(-1).float().clone()
        return st

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, grad_scale_t: Optional[torch.Tensor] = None):
        loss = closure() if closure is not None else None
        C = native()
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            flat = self._flat.get(gi)
            if flat is not None and flat[1] is not None and C is not None and flat[0].is_cuda:
                fp, fg = flat
                st = self._state_for(group["params"][0], fp)
                st["step"] += 1
                C.adamw(fp, st.get("master"), fg, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1,
                        b2, group["eps"], group["weight_decay"], st["step"], self.adamw_mode,
                        grad_scale, grad_scale_t, None)
                continue
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self._state_for(p, p)
                st["step"] += 1
                if C is not None and p.is_cuda and p.is_contiguous() and p.grad.is_contiguous() \
                        and p.dtype in (torch.bfloat16, torch.float32):
                    C.adamw(p.view(-1), st.get("master"), p.grad.view(-1), st["exp_avg"],
                            st["exp_avg_sq"], group["lr"], b1, b2, group["eps"],
                            group["weight_decay"], st["step"], self.adamw_mode, grad_scale,
                            grad_scale_t, None)
                else:
                    g = p.grad.float().reshape(-1) * grad_scale
                    w = st["master"] if "master" in st else p.data.float().reshape(-1)
                    if self.adamw_mode:
                        w.mul_(1 - group["lr"] * group["weight_decay"])
                    else:
                        g = g + group["weight_decay"] * w
                    st["exp_avg"].mul_(b1).add_(g, alpha=1 - b1)
                    st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
                    bc1 = 1 - b1 ** st["step"]
                    bc2 = 1 - b2 ** st["step"]
                    denom = (st["exp_avg_sq"].sqrt() / (bc2 ** 0.5)).add_(group["eps"])
                    w.addcdiv_(st["exp_avg"], denom, value=-group["lr"] / bc1)
                    p.data.copy_(w.view_as(p))
        return loss


class BucketAdamW:
    """AdamW that runs once per :class:`NaiveDDP` gradient bucket.

    NaiveDDP (``gradient_as_bucket_view=True``) keeps every gradient inside a few flat (symmetric
    memory) buckets that the NVLS all-reduce kernel averages in place.  This optimizer mirrors
    that layout for the parameters -- ``p.data`` becomes a view of a per-bucket flat parameter
    buffer at the same offset as its gradient -- so the whole update of a bucket (fp32 master
    weights + moments, bf16 write-out, optional clip coefficient) is ONE fused kernel launch that
    can start as soon as that bucket's reduction has finished.  ``zero_grad`` is one memset per
    bucket.  Exposes ``param_groups`` (one group) for LR schedulers.
    """

    def __init__(self, ddp, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 master_weights: bool = True, adamw_mode: bool = True,
                 fused_comm: Optional[bool] = None):
        """``fused_comm`` (default: on whenever the DDP buckets live in NVSwitch symmetric memory,
        ``TDP_FUSED_OPT=0`` turns it off): the data-parallel reduction and the optimizer become ONE
        kernel per bucket -- reduce-scatter in the switch, AdamW on this rank's 1/N shard of the
        fp32 master weights / moments, multicast all-gather of the new bf16 parameters
        (csrc/coll/collectives.cu ``fused_rs_adamw_ag_kernel``).  It runs on the DDP comm stream as
        soon as a bucket's gradients are complete, i.e. overlapped with the rest of backward;
        optimizer state and optimizer memory traffic shrink to 1/N per rank and ``step()`` has
        nothing left to do.  In this mode ``p.grad`` keeps the *local* gradient after the step
        (set ``write_back_grad = True`` to also publish the averaged one, e.g. for checks) and
        ``grad_scale`` / clipping between reduction and update is not available."""
        red = ddp.reducer
        if not red.as_view:
            raise ValueError("BucketAdamW needs NaiveDDP(gradient_as_bucket_view=True)")
        self.ddp = ddp
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                  params=[p for p in red.params.values() if p.requires_grad])]
        self.adamw_mode = adamw_mode
        self.step_count = 0
        self.state = []
        # device-resident {step, lr}: the kernel derives the bias corrections from it, so a
        # CUDA-graph replay of step() keeps advancing (host scalars are frozen at capture time)
        self._hyper = None
        self._hyper_lr = None
        if red.on_cuda:
            self._hyper = torch.tensor([0.0, float(lr)], dtype=torch.float32, device=red.device)
            self._hyper_lr = float(lr)
        import os
        world = red._group_size(red.default_group)
        can_fuse = (red.on_cuda and native() is not None and world > 1 and master_weights
                    and len(red.buckets) > 0
                    and all(b.symm is not None and b.dtype == torch.bfloat16
                            and red._group_size(b.group) == world for b in red.buckets))
        if fused_comm is None:
            fused_comm = can_fuse and os.environ.get("TDP_FUSED_OPT", "1") != "0"
        elif fused_comm and not can_fuse:
            raise ValueError("BucketAdamW(fused_comm=True) needs bf16 NaiveDDP buckets in symmetric "
                             "memory (NCCL group of 2..8 GPUs on one NVSwitch domain)")
        self.fused_comm = bool(fused_comm)
        self.write_back_grad = False
        self._step_open = False
        self._world = world
        param_sbufs = {}
        with torch.no_grad():
            for b in red.buckets:
                if self.fused_comm:
                    # parameters mirror the gradient layout inside a second symmetric buffer
                    gsbuf, goff = b.symm
                    if id(gsbuf) not in param_sbufs:
                        param_sbufs[id(gsbuf)] = gsbuf.group.alloc(gsbuf.nbytes)
                    psbuf = param_sbufs[id(gsbuf)]
                    flat_p = psbuf.view(goff, (b.capacity,), b.dtype)
                else:
                    flat_p = torch.zeros(b.capacity, dtype=b.dtype, device=b.device)
                esize = b.buffer.element_size()
                for name in b.names:
                    p = red.params[name]
                    off = (b.views[name].data_ptr() - b.buffer.data_ptr()) // esize
                    pv = flat_p[off:off + p.numel()].view(p.shape)
                    pv.copy_(p.data)
                    p.data = pv
                if self.fused_comm:
                    n = b.payload().numel()                       # multiple of 256 elements
                    n_vec = n // 8
                    per = (n_vec + world - 1) // world            # 16-byte vectors per rank
                    rank = dist_rank(red.default_group)
                    lo, hi = min(per * rank, n_vec) * 8, min(per * (rank + 1), n_vec) * 8
                    master = torch.zeros(per * 8, dtype=torch.float32, device=b.device)
                    master[:hi - lo].copy_(flat_p[lo:hi])
                    st = dict(bucket=b, flat_p=flat_p, master=master, numel=n, lo=lo, hi=hi,
                              exp_avg=torch.zeros_like(master), exp_avg_sq=torch.zeros_like(master),
                              param_sbuf=psbuf, off=goff)
                    b.fused_step = self._make_fused_step(st)
                else:
                    st = dict(bucket=b, flat_p=flat_p,
                              exp_avg=torch.zeros(b.capacity, dtype=torch.float32, device=b.device),
                              exp_avg_sq=torch.zeros(b.capacity, dtype=torch.float32, device=b.device),
                              master=(flat_p.float().clone() if master_weights and
                                      b.dtype != torch.float32 else None))
                self.state.append(st)
        if self.fused_comm:
            for psbuf in param_sbufs.values():
                psbuf.barrier(0)      # every rank's parameters are in place before anyone multicasts

    def _make_fused_step(self, st):
        def run(bucket):
            """Called by the DDP reducer on its comm stream when the bucket's gradients are final."""
            g = self.param_groups[0]
            if not self._step_open:
                self._hyper[0:1].add_(1.0)            # one optimizer step per backward pass
                self._step_open = True
            gsbuf, goff = bucket.symm
            b1, b2 = g["betas"]
            gsbuf.handle.fused_rs_adamw_ag(
                st["param_sbuf"].handle, goff, st["off"], st["numel"], st["master"], st["exp_avg"],
                st["exp_avg_sq"], float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                bool(self.adamw_mode), self._hyper,
                (1.0 / self._world) if self.ddp.reducer.average else 1.0,
                bool(self.write_back_grad), 0)
        return run

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, grad_scale_t: Optional[torch.Tensor] = None):
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        self.step_count += 1
        C = native()
        if self.fused_comm:
            # the update already happened, bucket by bucket, inside the reduction kernels that
            # NaiveDDP launched during backward / reduce_gradients()
            if grad_scale != 1.0 or grad_scale_t is not None:
                raise ValueError("BucketAdamW(fused_comm=True): no grad_scale / clipping hook "
                                 "between reduction and update; build it with fused_comm=False")
            if not self.ddp.reducer._finalized:
                self.ddp.reducer.finalize()
            if not torch.cuda.is_current_stream_capturing() and g["lr"] != self._hyper_lr:
                self._hyper[1].fill_(float(g["lr"]))
                self._hyper_lr = float(g["lr"])
            self._step_open = False
            return
        if self._hyper is not None:
            if not torch.cuda.is_current_stream_capturing() and g["lr"] != self._hyper_lr:
                self._hyper[1].fill_(float(g["lr"]))      # LR schedulers edit param_groups
                self._hyper_lr = float(g["lr"])
            self._hyper[0:1].add_(1.0)                    # captured: advances on every replay
        for st in self.state:
            fp, fg = st["flat_p"], st["bucket"].buffer[:st["bucket"].capacity]
            if C is not None and fp.is_cuda and fp.dtype in (torch.bfloat16, torch.float32):
                C.adamw(fp, st["master"], fg, st["exp_avg"], st["exp_avg_sq"], g["lr"], b1, b2,
                        g["eps"], g["weight_decay"], self.step_count, self.adamw_mode, grad_scale,
                        grad_scale_t, None, self._hyper)
            else:
                gr = fg.float() * grad_scale
                if grad_scale_t is not None:
                    gr = gr * grad_scale_t
                w = st["master"] if st["master"] is not None else fp.float()
                if self.adamw_mode:
                    w.mul_(1 - g["lr"] * g["weight_decay"])
                else:
                    gr = gr + g["weight_decay"] * w
                st["exp_avg"].mul_(b1).add_(gr, alpha=1 - b1)
                st["exp_avg_sq"].mul_(b2).addcmul_(gr, gr, value=1 - b2)
                bc1, bc2 = 1 - b1 ** self.step_count, 1 - b2 ** self.step_count
                denom = (st["exp_avg_sq"].sqrt() / (bc2 ** 0.5)).add_(g["eps"])
                w.addcdiv_(st["exp_avg"], denom, value=-g["lr"] / bc1)
                fp.copy_(w)

    def zero_grad(self, set_to_none: bool = False) -> None:
        for st in self.state:
            st["bucket"].buffer.zero_()

    def set_lr(self, lr: float) -> None:
        """Change the learning rate *outside* a captured step: the kernel reads lr from device
        memory, so the next CUDA-graph replay picks it up (editing ``param_groups`` alone is only
        seen by the Python ``step()``, which does not run during replay)."""
        self.param_groups[0]["lr"] = float(lr)
        if self._hyper is not None:
            self._hyper[1].fill_(float(lr))
            self._hyper_lr = float(lr)

    def current_step(self) -> int:
        """Number of optimizer steps taken, including CUDA-graph replays (device counter)."""
        if self._hyper is not None:
            return int(round(float(self._hyper[0].item())))
        return self.step_count

    def state_tensors(self) -> List[torch.Tensor]:
        """Every tensor ``step()`` mutates (for ``GraphedStep(preserve=...)``)."""
        out = [] if self._hyper is None else [self._hyper]
        for st in self.state:
            out += [st["flat_p"], st["exp_avg"], st["exp_avg_sq"]]
            if st["master"] is not None:
                out.append(st["master"])
        return out

    def state_dict(self) -> dict:
        self.step_count = self.current_step()      # replays advance only the device counter
        return dict(step=self.step_count, fused_comm=self.fused_comm, world=self._world, param_groups=[{k: v for k, v in g.items() if k != "params"}
                                                        for g in self.param_groups],
                    buckets=[dict(exp_avg=s["exp_avg"].cpu(), exp_avg_sq=s["exp_avg_sq"].cpu(),
                                  master=None if s["master"] is None else s["master"].cpu())
                             for s in self.state])

    def load_state_dict(self, sd: dict) -> None:
        if bool(sd.get("fused_comm", False)) != self.fused_comm or \
                (self.fused_comm and int(sd.get("world", self._world)) != self._world):
            raise ValueError("BucketAdamW checkpoint was written with a different fused_comm / "
                             "world size (optimizer shards are rank-local in fused mode)")
        self.step_count = int(sd["step"])
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            g.update(sg)
        if self._hyper is not None:            # the device-resident {step, lr} the kernel reads
            self._hyper[0].fill_(float(self.step_count))
            self._hyper[1].fill_(float(self.param_groups[0]["lr"]))
            self._hyper_lr = float(self.param_groups[0]["lr"])
        for s, ss in zip(self.state, sd["buckets"]):
            s["exp_avg"].copy_(ss["exp_avg"])
            s["exp_avg_sq"].copy_(ss["exp_avg_sq"])
            if s["master"] is not None and ss["master"] is not None:
                s["master"].copy_(ss["master"])
                if self.fused_comm:
                    # my slice of the bf16 parameters; the other slices come from the model
                    # checkpoint (identical on every rank)
                    s["flat_p"][s["lo"]:s["hi"]].copy_(s["master"][:s["hi"] - s["lo"]])
                else:
                    s["flat_p"].copy_(s["master"])


def flatten_module_params(module: torch.nn.Module, align_elems: int = 64):
    """Re-home all parameters of ``module`` (single dtype) into one flat buffer.  Returns
    ``flat_param``; each ``p.data`` becomes a view of it (registration order)."""
    params = [p for p in module.parameters()]
    assert params and all(p.dtype == params[0].dtype for p in params)
    offs, total = [], 0
    for p in params:
        total = (total + align_elems - 1) // align_elems * align_elems
        offs.append(total)
        total += p.numel()
    total = (total + align_elems - 1) // align_elems * align_elems
    flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
    with torch.no_grad():
        for p, o in zip(params, offs):
            v = flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
    return flat, offs


class TensorListTable:
    """Device-side pointer / numel / dtype tables of a list of (contiguous bf16 / fp32) tensors,
    for the multi-tensor kernels (one launch for the whole list).  Tables are cached by the
    tensors' addresses -- gradient lists of a training loop are stable (bucket views)."""

    _cache: dict = {}

    def __init__(self, tensors: List[torch.Tensor]):
        dev = tensors[0].device
        self.keep = list(tensors)
        self.ptrs = torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device=dev)
        self.numels = torch.tensor([t.numel() for t in tensors], dtype=torch.int64, device=dev)
        self.dtypes = torch.tensor([0 if t.dtype == torch.bfloat16 else 1 for t in tensors],
                                   dtype=torch.int32, device=dev)

    @classmethod
    def of(cls, tensors: List[torch.Tensor]) -> "TensorListTable":
        key = tuple((t.data_ptr(), t.numel(), t.dtype) for t in tensors)
        tab = cls._cache.get(key)
        if tab is None:
            if len(cls._cache) > 64:
                cls._cache.clear()
            tab = cls._cache[key] = cls(tensors)
        return tab


def _multi_ok(t: torch.Tensor) -> bool:
    return t.is_cuda and t.is_contiguous() and t.dtype in (torch.bfloat16, torch.float32)


def multi_sumsq(tensors: List[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum_i ||t_i||^2 as a 1-element fp32 device tensor: ONE kernel launch for all supported
    tensors (csrc/fused/optim.cu ``sumsq_multi_kernel``), torch fallback for the rest."""
    tensors = [t for t in tensors if t.numel() > 0]
    dev = tensors[0].device if tensors else torch.device("cpu")
    acc = out if out is not None else torch.zeros(1, dtype=torch.float32, device=dev)
    C = native() if dev.type == "cuda" else None
    fast = [t for t in tensors if C is not None and _multi_ok(t)]
    if fast:
        tab = TensorListTable.of(fast)
        C.sumsq_multi(tab.ptrs, tab.numels, tab.dtypes, acc)
    for t in tensors:
        if C is None or not _multi_ok(t):
            acc += t.detach().float().pow(2).sum()
    return acc


def multi_scale_(tensors: List[torch.Tensor], scale: float = 1.0,
                 scale_t: Optional[torch.Tensor] = None) -> None:
    """t_i *= scale (* scale_t[0], a device scalar -- no host sync) for every tensor, one launch."""
    tensors = [t for t in tensors if t.numel() > 0]
    if not tensors:
        return
    dev = tensors[0].device
    C = native() if dev.type == "cuda" else None
    fast = [t for t in tensors if C is not None and _multi_ok(t)]
    if fast:
        tab = TensorListTable.of(fast)
        C.scale_multi(tab.ptrs, tab.numels, tab.dtypes, float(scale),
                      None if scale_t is None else scale_t.reshape(1).float())
    for t in tensors:
        if C is None or not _multi_ok(t):
            f = scale if scale_t is None else scale_t.to(t.dtype) * scale
            t.detach().mul_(f)
