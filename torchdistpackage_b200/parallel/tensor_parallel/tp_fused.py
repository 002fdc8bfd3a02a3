"""Sequence-parallel linear layers whose collective is fused into the GEMM kernel
(NVSwitch peer memory, no NCCL on this path).

    ag_linear(ctx, x_shard, W, b)        y_full  = all_gather(x_shard) @ W + b
    linear_rs(ctx, a_full, W, b)         y_shard = reduce_scatter(a_full @ W) + b
    sp_mlp(ctx, x_shard, W1,b1, W2,b2)   y_shard = RS(gelu(AG(x) @ W1 + b1) @ W2) + b2

Mechanics (csrc/gemm/gemm_sm100.cuh, csrc/coll/collectives.cu):

* all-gather -> GEMM: a push kernel on a side stream multicasts this rank's shard into every
  rank's gather buffer (``multimem.st``) and then raises one flag per source chunk; the GEMM's
  TMA producer warp starts on the local chunk (read zero-copy from the shard itself), then
  consumes chunks ``rank+1, rank+2, ...`` as their flags arrive -- tensor-core work on chunk i
  overlaps the NVLink transfer of chunk i+1.
* GEMM -> reduce-scatter: tiles are computed remote-chunks-first; each finished 128x64 sub-tile
  is TMA-stored straight into the owner's staging slot and a counter is bumped with release
  semantics; the owner's reduce kernel waits on the counters, sums the ``tp`` partials in fp32 and
  adds the bias (+ residual).
* backward mirrors forward: d(AG->GEMM) is a GEMM->RS, d(GEMM->RS) is an AG->GEMM (with the GELU
  derivative in its epilogue for the MLP); weight gradients use the gathered activations that are
  still sitting in the symmetric gather buffers (one buffer per module and direction, sized for
  180 GB parts -- nothing is re-gathered, nothing is copied out).

Replaces: ``gather_from_sequence_parallel_region`` + ``torch.matmul`` and ``torch.matmul`` +
``reduce_scatter_to_sequence_parallel_region`` of the reference (tp_utils.py:52-159,
attn.py:93-98, mlp.py:69-78), which run back to back on one stream.
"""
from __future__ import annotations

import os

from typing import Dict, Tuple

import torch
import torch.distributed as dist

from ...ops import linear as L
from ...ops._loader import native
from ...ops.symm import SymmBuffer, get_symm_group

_AG_PUSH_CTAS = 16      # side-stream push variant: SMs left free by the GEMM for the push CTAs
_AG_PUSH_IN_KERNEL = os.environ.get("TDP_AG_PUSH", "kernel") != "side"
_ENABLED = True


def set_enabled(flag: bool) -> None:
    """Globally switch the fused paths off (falls back to NCCL collectives + plain GEMMs)."""
    global _ENABLED
    _ENABLED = bool(flag)


def usable(x: torch.Tensor, tp_size: int) -> bool:
    if not _ENABLED or tp_size <= 1 or not x.is_cuda or x.dtype != torch.bfloat16:
        return False
    if native() is None:
        return False
    rows = x.numel() // x.shape[-1]
    return rows % 128 == 0 and x.shape[-1] % 8 == 0


class _Region:
    """A symmetric buffer + its flag words + side-stream bookkeeping."""

    def __init__(self, sg, nbytes: int):
        # Ping-pong halves: use k writes half k%2.  Re-using a half two uses later needs NO
        # cross-GPU barrier: before rank A starts use k+2 it has finished use k+1, which consumed
        # every peer's data of use k+1, and a peer only produces use k+1 after it has finished
        # consuming use k (stream order) -- so all readers of half k%2 are done.
        self.half = (int(nbytes) + 4095) // 4096 * 4096
        self.buf: SymmBuffer = sg.alloc(2 * self.half)
        self.flag_word = self.buf.alloc_words(8)     # chunk flags (AG) / tile counters (RS)
        self.nbytes = nbytes
        self.uses = 0

    def next_offset(self) -> int:
        off = (self.uses & 1) * self.half
        self.uses += 1
        return off


class FusedSpContext:
    """Per-module workspace: named symmetric regions, allocated lazily (collectively, in the same
    order on every rank) and re-used every step."""

    def __init__(self, group):
        self.group = group
        self.sg = get_symm_group(group)
        if not self.sg.enabled:
            raise RuntimeError(f"fused sequence-parallel kernels need symmetric memory: "
                               f"{self.sg.reason}")
        self.tp = self.sg.world
        self.rank = self.sg.rank
        self.regions: Dict[Tuple[str, int], _Region] = {}
        _, hi = torch.cuda.Stream.priority_range()
        self.side = torch.cuda.Stream(priority=hi)

    def region(self, name: str, nbytes: int) -> _Region:
        key = (name, nbytes)
        r = self.regions.get(key)
        if r is None:
            r = _Region(self.sg, nbytes)
            self.regions[key] = r
        return r


# ------------------------------------------------------------------------------------------
# building blocks (no autograd)
# ------------------------------------------------------------------------------------------
def _ag_gemm(ctx: FusedSpContext, name: str, x_shard: torch.Tensor, w: torch.Tensor,
             trans_b: bool, bias=None, act: int = 0, aux_in=None, want_aux_out: bool = False):
    """Returns (out_full [T, N], gathered [T, K] view of the symmetric buffer, aux_out or None)."""
    rows, K = x_shard.shape
    T = rows * ctx.tp
    N = w.shape[0] if trans_b else w.shape[1]
    reg = ctx.region(name, T * K * 2)
    buf = reg.buf
    out = torch.empty(T, N, dtype=torch.bfloat16, device=x_shard.device)
    aux_out = torch.empty(T, N, dtype=torch.bfloat16, device=x_shard.device) if want_aux_out else None
    off = reg.next_offset()                  # ping-pong half: no barrier needed (see _Region)
    ep = buf.next_epoch(reg.flag_word)
    if _AG_PUSH_IN_KERNEL and x_shard.is_contiguous():
        # one kernel: an extra warp of every GEMM CTA pushes its share of the local shard to all
        # peers (multimem.st) while the tensor cores start on the local chunk
        buf.handle.gemm_ag(off, rows, K, w, trans_b, out, bias, aux_out, act, reg.flag_word, ep,
                           0, x_shard, aux_in, True, True)
    else:
        cur = torch.cuda.current_stream()
        ctx.side.wait_stream(cur)
        with torch.cuda.stream(ctx.side):
            buf.handle.all_gather_signal(off, rows * K * 2, x_shard, reg.flag_word, ep, True,
                                         _AG_PUSH_CTAS)
        x_shard.record_stream(ctx.side)
        buf.handle.gemm_ag(off, rows, K, w, trans_b, out, bias, aux_out, act, reg.flag_word, ep,
                           L._num_sms() - _AG_PUSH_CTAS, x_shard, aux_in)
        cur.wait_stream(ctx.side)
    gathered = buf.view(off, (T, K), torch.bfloat16)
    gathered._tdp_token = (reg, ep)          # lets backward detect that the buffer was re-used
    return out, gathered, aux_out


def _gathered_or_regather(ctx: FusedSpContext, gathered: torch.Tensor, token, shard: torch.Tensor):
    """The gathered activation saved by forward lives in a symmetric buffer that the next forward
    of the same module overwrites (several micro-batches in flight, activation checkpointing).
    If that happened, rebuild it from the saved shard."""
    reg, ep = token
    # ping-pong halves: the data survives exactly one later use of the region
    if ((reg.buf._epochs.get(reg.flag_word, 0) - ep) & 0xFFFFFFFF) <= 1:
        return gathered
    full = torch.empty_like(gathered)
    dist.all_gather_into_tensor(full, shard.contiguous(), group=ctx.group)
    return full


def _gemm_rs(ctx: FusedSpContext, name: str, a: torch.Tensor, w: torch.Tensor, trans_b: bool,
             bias=None, residual=None) -> torch.Tensor:
    """out_shard [T/tp, N] = reduce_scatter(a @ op(w)) + bias (+ residual)."""
    T, _ = a.shape
    rows = T // ctx.tp
    N = w.shape[0] if trans_b else w.shape[1]
    reg = ctx.region(name, T * N * 2)
    buf = reg.buf
    out = torch.empty(rows, N, dtype=torch.bfloat16, device=a.device)
    off = reg.next_offset()                  # ping-pong staging: no barrier needed (see _Region)
    buf.handle.gemm_rs(a, w, trans_b, off, reg.flag_word, 0)
    target = buf.next_epoch(reg.flag_word, (rows // 32) * (N // 8))
    buf.handle.rs_reduce(off, rows, N, reg.flag_word, target, bias, residual, out, False, 0, True, 0)
    return out


def _gemm_ar(ctx: FusedSpContext, name: str, a: torch.Tensor, w: torch.Tensor, trans_b: bool,
             bias=None) -> torch.Tensor:
    """out_full [T, N] = all_reduce(a @ op(w)) + bias, fused: the GEMM epilogue scatters tiles to
    the chunk owners (as in GEMM->RS); each owner's reduce kernel then *broadcasts* its reduced
    rows to every rank with ``multimem.st`` (two-shot all-reduce whose first shot is the GEMM
    epilogue).  The result is copied out of the (reused) symmetric output buffer: autograd may
    keep it across several forwards of the same module (1F1B warm-up, forward-all-then-backward
    micro-batching), during which the buffer is overwritten."""
    T, _ = a.shape
    rows = T // ctx.tp
    N = w.shape[0] if trans_b else w.shape[1]
    nb = T * N * 2
    reg = ctx.region(name, 2 * nb)           # [staging | output] per ping-pong half
    buf = reg.buf
    off = reg.next_offset()
    buf.handle.gemm_rs(a, w, trans_b, off, reg.flag_word, 0)
    target = buf.next_epoch(reg.flag_word, (rows // 32) * (N // 8))
    buf.handle.rs_reduce(off, rows, N, reg.flag_word, target, bias, None, None, True, off + nb,
                         True, 0)
    buf.barrier(0)                           # every owner's broadcast has landed everywhere
    return buf.view(off + nb, (T, N), torch.bfloat16).clone()


class _LinearArFn(torch.autograd.Function):
    """Row-parallel linear without sequence parallelism: y = all_reduce(a_local @ W_local) + b."""

    @staticmethod
    def forward(ctx, fctx, slot, a, w, bias):
        out = _gemm_ar(fctx, slot + ":ar", a, w, False, bias)
        ctx.save_for_backward(a, w)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        a, w = ctx.saved_tensors
        dy = dy.contiguous()
        da = L.gemm(dy, w, trans_b=True)          # the all-reduce is an identity in backward
        dw = L.gemm(a, dy, trans_a=True)
        db = L.colsum(dy) if ctx.has_bias else None
        return None, None, da, dw, db


def linear_ar(fctx: FusedSpContext, a: torch.Tensor, w, bias, slot: str = "out"):
    return _LinearArFn.apply(fctx, slot, a.contiguous(), w, bias)


# ------------------------------------------------------------------------------------------
# autograd functions
# ------------------------------------------------------------------------------------------
class _AgLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fctx, slot, x_shard, w, bias):
        out, gathered, _ = _ag_gemm(fctx, slot + ":fwd", x_shard, w, False, bias)
        ctx.fctx, ctx.slot = fctx, slot
        ctx.token = gathered._tdp_token
        ctx.save_for_backward(gathered, w, x_shard)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x_full, w, x_shard = ctx.saved_tensors
        x_full = _gathered_or_regather(ctx.fctx, x_full, ctx.token, x_shard)
        dy = dy.contiguous()
        dw = L.gemm(x_full, dy, trans_a=True)                 # [K, N] = x_full^T @ dy
        db = L.colsum(dy) if ctx.has_bias else None
        dx = _gemm_rs(ctx.fctx, ctx.slot + ":bwd", dy, w, True)   # RS(dy @ W^T)
        return None, None, dx, dw, db


class _LinearRsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fctx, slot, a, w, bias):
        out = _gemm_rs(fctx, slot + ":fwd", a, w, False, bias)
        ctx.fctx, ctx.slot = fctx, slot
        ctx.save_for_backward(a, w)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dy_shard):
        a, w = ctx.saved_tensors
        dy_shard = dy_shard.contiguous()
        da, dy_full, _ = _ag_gemm(ctx.fctx, ctx.slot + ":bwd", dy_shard, w, True)  # AG(dy) @ W^T
        dw = L.gemm(a, dy_full, trans_a=True)
        db = L.colsum(dy_shard) if ctx.has_bias else None     # partial over tp (sequence shard)
        return None, None, da, dw, db


class _SpMlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fctx, x_shard, w1, b1, w2, b2, act: int):
        a, x_full, z = _ag_gemm(fctx, "mlp_in:fwd", x_shard, w1, False, b1, act, None, True)
        y = _gemm_rs(fctx, "mlp_out:fwd", a, w2, False, b2)
        ctx.fctx, ctx.act = fctx, act
        ctx.token = x_full._tdp_token
        ctx.save_for_backward(x_full, w1, w2, z, a, x_shard)
        ctx.flags = (b1 is not None, b2 is not None)
        return y

    @staticmethod
    def backward(ctx, dy_shard):
        x_full, w1, w2, z, a, x_shard = ctx.saved_tensors
        fctx = ctx.fctx
        x_full = _gathered_or_regather(fctx, x_full, ctx.token, x_shard)
        dy_shard = dy_shard.contiguous()
        # dz = (AG(dy) @ W2^T) * gelu'(z): derivative applied in the AG-GEMM epilogue
        dz, dy_full, _ = _ag_gemm(fctx, "mlp_out:bwd", dy_shard, w2, True, None,
                                  L._DACT[ctx.act], z)
        dw2 = L.gemm(a, dy_full, trans_a=True)
        db2 = L.colsum(dy_shard) if ctx.flags[1] else None
        dx = _gemm_rs(fctx, "mlp_in:bwd", dz, w1, True)
        dw1 = L.gemm(x_full, dz, trans_a=True)
        db1 = L.colsum(dz) if ctx.flags[0] else None
        return None, dx, dw1, db1, dw2, db2, None


def ag_linear(fctx: FusedSpContext, x_shard: torch.Tensor, w, bias, slot: str = "in"):
    return _AgLinearFn.apply(fctx, slot, x_shard.contiguous(), w, bias)


def linear_rs(fctx: FusedSpContext, a: torch.Tensor, w, bias, slot: str = "out"):
    return _LinearRsFn.apply(fctx, slot, a.contiguous(), w, bias)


def sp_mlp(fctx: FusedSpContext, x_shard: torch.Tensor, w1, b1, w2, b2, act: str = "gelu"):
    return _SpMlpFn.apply(fctx, x_shard.contiguous(), w1, b1, w2, b2, L._ACT_CODE[act])
