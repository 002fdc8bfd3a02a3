"""GPT-2 as pipeline stages of tensor(+sequence)-parallel blocks (BASELINE.json config #5:
GPT-2 medium, DP=2 x PP=2 x TP=2 with ZeRO over the data group).

    stage = GPT2PipelineStage(cfg, tp_group=tpc.get_group('tensor'))      # layers of *this* pp rank
    loss  = forward_backward(zero_opt, stage.forward_fn(), None, stage.stage_inputs(tokens, targets),
                             num_microbatches=m, dtype=torch.bfloat16)
    stage.allreduce_replicated_grads(); zero_opt.step()

Layout: activations between blocks are sequence-parallel shards ``[B/tp, T, D]`` (dim 0 split, the
reference's convention), so the pipeline p2p messages are already 1/tp of the activation.  The
first stage owns the embeddings, the last stage the final LayerNorm and an (untied) LM head; the
LM head works on the local shard and the loss is averaged over the tensor group.  Parameters that
are replicated over the tensor group but only see 1/tp of the tokens (LayerNorms, row-parallel
biases, embeddings, head) are tagged ``sequence_parallel_grad`` and summed over the tensor group
by ``allreduce_replicated_grads``.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from ..dist.process_topo import tpc
from ..ops import fused as F_ops
from ..parallel.pipeline_parallel.pipeline_helper import uniform_bounds
from ..parallel.tensor_parallel.tp_utils import (set_tp_group, _split_along_first_dim,
                                                  set_sequence_parallel_attr)
from ..parallel.tensor_parallel.transformer import ParallelBlock, allreduce_sequence_parallel_grads
from .gpt2 import GPT2Config


class GPT2PipelineStage(nn.Module):
    def __init__(self, cfg: GPT2Config, tp_group=None, sequence_parallel: bool = True,
                 pp_rank: Optional[int] = None, pp_size: Optional[int] = None):
        super().__init__()
        self.cfg = cfg
        self.tp_group = tp_group
        set_tp_group(tp_group)
        self.tp = dist.get_world_size(tp_group) if dist.is_initialized() else 1
        self.sp = sequence_parallel and self.tp > 1
        self.pp_size = pp_size if pp_size is not None else (
            tpc.get_group_size("pipe") if tpc.is_mode_inited("pipe") else 1)
        self.pp_rank = pp_rank if pp_rank is not None else (
            tpc.get_group_rank("pipe") if tpc.is_mode_inited("pipe") else 0)
        self.first, self.last = self.pp_rank == 0, self.pp_rank == self.pp_size - 1
        beg, end = uniform_bounds(cfg.n_layer, self.pp_size)[self.pp_rank]
        self.layer_range = (beg, end)
        if self.first:
            self.wte = nn.Embedding(cfg.vocab_size, cfg.d_model)
            self.wpe = nn.Embedding(cfg.seq_len, cfg.d_model)
            nn.init.normal_(self.wte.weight, std=0.02)
            nn.init.normal_(self.wpe.weight, std=0.02)
        self.blocks = nn.ModuleList([
            ParallelBlock(cfg.d_model, mlp_ratio=cfg.mlp_ratio, num_heads=cfg.n_head,
                          sequence_parallel=self.sp, causal=True) for _ in range(end - beg)])
        for blk in self.blocks:
            for m in blk.modules():
                if hasattr(m, "reset_parameters_scaled"):
                    m.reset_parameters_scaled()
        if self.last:
            self.ln_f = nn.LayerNorm(cfg.d_model)
            self.lm_head = nn.Parameter(torch.empty(cfg.vocab_size, cfg.d_model))
            nn.init.normal_(self.lm_head, std=0.02)
        if self.sp:
            tagged = []
            if self.first:
                tagged += [self.wte.weight, self.wpe.weight]
            if self.last:
                tagged += list(self.ln_f.parameters()) + [self.lm_head]
            for p in tagged:
                p.sequence_parallel_grad = True

    # ---------------------------------------------------------------- per-micro-batch forward
    def forward(self, x: torch.Tensor, tokens: Optional[torch.Tensor] = None,
                targets: Optional[torch.Tensor] = None):
        if self.first:
            T = tokens.shape[1]
            h = self.wte(tokens) + self.wpe(torch.arange(T, device=tokens.device))
            x = _split_along_first_dim(h) if self.sp else h          # [B/tp, T, D]
        elif self.sp:
            set_sequence_parallel_attr(x)       # what arrives over the pipe is already a shard
        for blk in self.blocks:
            x = blk(x)
        if not self.last:
            return x
        x = F_ops.layer_norm(x, self.ln_f.weight, self.ln_f.bias, self.ln_f.eps)
        if self.sp:
            k = targets.shape[0] // self.tp
            r = dist.get_rank(self.tp_group)
            targets = targets[r * k:(r + 1) * k]
        loss = F_ops.lm_head_loss(x, self.lm_head, targets)
        return loss / self.tp if self.sp else loss       # summed over the tensor group = mean

    def forward_fn(self):
        """The ``fwd_fn`` for ``forward_backward``: unpacks [prev_activation] + this stage's own
        micro-batch inputs (tokens on the first stage, targets on the last)."""
        def fn(inp):
            items = list(inp) if isinstance(inp, (list, tuple)) else [inp]
            x = None if self.first else items.pop(0)
            tokens = items.pop(0) if self.first else None
            targets = items.pop(0) if self.last else None
            return self.forward(x, tokens, targets)
        return fn

    def stage_inputs(self, tokens: torch.Tensor, targets: torch.Tensor) -> Optional[List[torch.Tensor]]:
        ins = ([tokens] if self.first else []) + ([targets] if self.last else [])
        return ins or None

    def allreduce_replicated_grads(self) -> None:
        if self.sp:
            allreduce_sequence_parallel_grads(self, self.tp_group)
