"""MoE transformer LM (BASELINE.json config #4: 8 experts, EP=4 x moe-DP=2 on 8 GPUs).

GPT-2 style blocks whose MLP is a :class:`~torchdistpackage_b200.moe.layer.MoELayer` (every
``moe_every``-th block).  Attention / embeddings / LayerNorm are replicated (reduced over the full
``data`` group by NaiveDDP); expert weights are sharded over ``moe_ep`` and replicated over
``moe_dp`` (reduced by ``create_moe_dp_hooks``).  ``ddp_ignore_names()`` returns the expert
parameter names for ``module._ddp_params_and_buffers_to_ignore``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn as nn

from ..moe.layer import MoELayer
from ..ops import fused as F_ops
from ..ops import linear as L_ops
from ..ops.attention import packed_attention
from .gpt2 import GPT2Block, GPT2Config


@dataclass
class MoEConfig(GPT2Config):
    n_layer: int = 4
    n_head: int = 16
    d_model: int = 1024
    num_experts: int = 8
    top_k: int = 2
    capacity_factor: float = 1.25
    moe_every: int = 1
    aux_loss_coef: float = 0.01

    @staticmethod
    def tiny() -> "MoEConfig":
        return MoEConfig(vocab_size=512, n_layer=2, n_head=4, d_model=128, seq_len=128,
                         num_experts=4)


class MoEBlock(nn.Module):
    def __init__(self, cfg: MoEConfig, ep_group=None):
        super().__init__()
        d = cfg.d_model
        self.n_head = cfg.n_head
        self.ln_1, self.ln_2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.w_qkv = nn.Parameter(torch.empty(d, 3 * d)); self.b_qkv = nn.Parameter(torch.zeros(3 * d))
        self.w_proj = nn.Parameter(torch.empty(d, d)); self.b_proj = nn.Parameter(torch.zeros(d))
        nn.init.normal_(self.w_qkv, std=0.02)
        nn.init.normal_(self.w_proj, std=0.02 / (2 * cfg.n_layer) ** 0.5)
        self.moe = MoELayer(d, cfg.mlp_ratio * d, cfg.num_experts, cfg.top_k, cfg.capacity_factor,
                            ep_group=ep_group)

    def forward(self, x):
        B, T, D = x.shape
        h = F_ops.layer_norm(x, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps)
        qkv = L_ops.linear(h, self.w_qkv, self.b_qkv, layout="kn")
        o = packed_attention(qkv, self.n_head, causal=True)
        x = L_ops.linear(o, self.w_proj, self.b_proj, layout="kn", residual=x)
        h = F_ops.layer_norm(x, self.ln_2.weight, self.ln_2.bias, self.ln_2.eps)
        y, aux = self.moe(h)
        return x + y, aux


class MoETransformer(nn.Module):
    def __init__(self, cfg: MoEConfig, ep_group=None):
        super().__init__()
        self.cfg = cfg
        self.wte = nn.Embedding(cfg.vocab_size, cfg.d_model)
        self.wpe = nn.Embedding(cfg.seq_len, cfg.d_model)
        self.blocks = nn.ModuleList([
            MoEBlock(cfg, ep_group) if (i % cfg.moe_every == 0) else GPT2Block(cfg)
            for i in range(cfg.n_layer)])
        self.ln_f = nn.LayerNorm(cfg.d_model)
        nn.init.normal_(self.wte.weight, std=0.02)
        nn.init.normal_(self.wpe.weight, std=0.02)

    def expert_parameters(self) -> Dict[str, nn.Parameter]:
        return {n: p for n, p in self.named_parameters() if getattr(p, "is_expert", False)}

    def ddp_ignore_names(self):
        return list(self.expert_parameters().keys())

    def forward(self, idx, targets: Optional[torch.Tensor] = None):
        x = self.wte(idx) + self.wpe(torch.arange(idx.shape[1], device=idx.device))
        aux_total = 0.0
        for blk in self.blocks:
            out = blk(x)
            if isinstance(out, tuple):
                x, aux = out
                aux_total = aux_total + aux
            else:
                x = out
        x = F_ops.layer_norm(x, self.ln_f.weight, self.ln_f.bias, self.ln_f.eps)
        logits = L_ops.linear(x, self.wte.weight, None, layout="nk")
        if targets is None:
            return logits
        loss = F_ops.cross_entropy(logits, targets)
        return loss + self.cfg.aux_loss_coef * aux_total
