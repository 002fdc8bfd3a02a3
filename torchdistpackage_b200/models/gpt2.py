"""GPT-2 (small / medium / ...) built on this package's sm_100a operators.

The reference ships no GPT model (its ``Transformer`` has no embeddings, mask, final LN or LM
head -- parallel/tensor_parallel/transformer.py:88-99); BASELINE.json's flagship configs name
GPT-2 small / medium, so the model lives here and is shared by ``bench.py``, the tests and the
examples.  Architecture = GPT-2: learned positional embeddings, pre-LN blocks, tanh-GELU MLP,
causal attention, tied input/output embedding, vocab padded to 50304.

Hot path per block (all bf16, fp32 accumulate):
    LN1 (fused kernel) -> qkv GEMM (+bias epilogue) -> causal flash attention (SDPA, library)
    -> proj GEMM (+bias +residual epilogue) -> LN2 -> fc1 GEMM (+bias +GELU epilogue, keeps z)
    -> fc2 GEMM (+bias +residual epilogue); backward applies GELU' in the fc2-dgrad epilogue.
Loss: LM-head GEMM against the tied embedding, then ONE fused softmax-cross-entropy kernel that
writes d(logits) in place (ops/fused.py).

Tensor / pipeline parallel variants reuse ``ParallelBlock`` (tensor_parallel/transformer.py) and
``flatten_model`` / ``partition_uniform`` (pipeline_parallel/pipeline_helper.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from ..ops import fused as F_ops
from ..ops import linear as L_ops
from ..ops.attention import packed_attention


@dataclass
class GPT2Config:
    vocab_size: int = 50304
    n_layer: int = 12
    n_head: int = 12
    d_model: int = 768
    seq_len: int = 1024
    mlp_ratio: int = 4

    @staticmethod
    def small() -> "GPT2Config":
        return GPT2Config()

    @staticmethod
    def medium() -> "GPT2Config":
        return GPT2Config(n_layer=24, n_head=16, d_model=1024)

    @staticmethod
    def tiny() -> "GPT2Config":   # unit tests / smoke
        return GPT2Config(vocab_size=512, n_layer=2, n_head=4, d_model=128, seq_len=128)

    def num_params(self) -> int:
        d, L, V = self.d_model, self.n_layer, self.vocab_size
        per_block = 4 * d + (3 * d * d + 3 * d) + (d * d + d) + 2 * (self.mlp_ratio * d * d) + \
            self.mlp_ratio * d + d
        return V * d + self.seq_len * d + L * per_block + 2 * d

    def flops_per_token(self) -> float:
        """fwd+bwd model FLOPs per token (6 * matmul params + attention)."""
        d, L = self.d_model, self.n_layer
        matmul_params = L * (4 * d * d + 2 * self.mlp_ratio * d * d) + self.vocab_size * d
        attn = L * 2 * 2 * self.seq_len * d / 2      # causal: half of QK^T and PV
        return 6.0 * matmul_params + 3.0 * attn


class GPT2Block(nn.Module):
    def __init__(self, cfg: GPT2Config):
        super().__init__()
        d, h = cfg.d_model, cfg.mlp_ratio * cfg.d_model
        self.n_head = cfg.n_head
        self.ln_1 = nn.LayerNorm(d)
        self.ln_2 = nn.LayerNorm(d)
        # weights are [in, out] (x @ W), the reference's TpLinear convention
        self.w_qkv = nn.Parameter(torch.empty(d, 3 * d))
        self.b_qkv = nn.Parameter(torch.zeros(3 * d))
        self.w_proj = nn.Parameter(torch.empty(d, d))
        self.b_proj = nn.Parameter(torch.zeros(d))
        self.w_fc1 = nn.Parameter(torch.empty(d, h))
        self.b_fc1 = nn.Parameter(torch.zeros(h))
        self.w_fc2 = nn.Parameter(torch.empty(h, d))
        self.b_fc2 = nn.Parameter(torch.zeros(d))
        std = 0.02
        for w in (self.w_qkv, self.w_fc1):
            nn.init.normal_(w, std=std)
        for w in (self.w_proj, self.w_fc2):
            nn.init.normal_(w, std=std / math.sqrt(2 * cfg.n_layer))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, T, D = x.shape
        # layer_norm_fork: LN(x) plus an alias of x for the skip connection, so the two gradients
        # of x are summed inside the LayerNorm backward kernel (no separate add kernels)
        h, skip = F_ops.layer_norm_fork(x, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps)
        qkv = L_ops.linear(h, self.w_qkv, self.b_qkv, layout="kn")
        o = packed_attention(qkv, self.n_head, causal=True)                # [B, T, D], no layout copies
        x = L_ops.linear(o, self.w_proj, self.b_proj, layout="kn", residual=skip)  # x + proj(o)
        h, skip = F_ops.layer_norm_fork(x, self.ln_2.weight, self.ln_2.bias, self.ln_2.eps)
        return L_ops.mlp(h, self.w_fc1, self.b_fc1, self.w_fc2, self.b_fc2, layout="kn",
                         act="gelu_tanh", residual=skip)                   # x + mlp(h)


class GPT2(nn.Module):
    """``forward(idx, targets)`` returns the mean token cross entropy (or logits if no targets)."""

    def __init__(self, cfg: GPT2Config):
        super().__init__()
        self.cfg = cfg
        self.wte = nn.Embedding(cfg.vocab_size, cfg.d_model)
        self.wpe = nn.Embedding(cfg.seq_len, cfg.d_model)
        self.blocks = nn.ModuleList([GPT2Block(cfg) for _ in range(cfg.n_layer)])
        self.ln_f = nn.LayerNorm(cfg.d_model)
        nn.init.normal_(self.wte.weight, std=0.02)
        # tied with the LM head: two gradient contributions per step, autograd must sum them
        self.wte.weight._tdp_no_fused_wgrad = True
        nn.init.normal_(self.wpe.weight, std=0.02)

    def embed(self, idx: torch.Tensor) -> torch.Tensor:
        T = idx.shape[1]
        # tied_embedding: the token-embedding gradient is scattered into the LM head's dense
        # gradient (ops/fused.py); positions are a plain slice (its backward is a dense copy,
        # not an index scatter)
        return F_ops.tied_embedding(idx, self.wte.weight) + self.wpe.weight[:T]

    def head_loss(self, x: torch.Tensor, targets: Optional[torch.Tensor]):
        x = F_ops.layer_norm(x, self.ln_f.weight, self.ln_f.bias, self.ln_f.eps)
        if targets is None:
            return L_ops.linear(x, self.wte.weight, None, layout="nk")     # tied embedding
        # LM head + CE + both LM-head backward GEMMs fused into one terminal op
        return F_ops.lm_head_loss(x, self.wte.weight, targets)

    def forward(self, idx: torch.Tensor, targets: Optional[torch.Tensor] = None):
        x = self.embed(idx)
        for blk in self.blocks:
            x = blk(x)
        return self.head_loss(x, targets)


def build_gpt2(name: str = "small", device=None, dtype=torch.bfloat16) -> GPT2:
    cfg = {"small": GPT2Config.small, "medium": GPT2Config.medium, "tiny": GPT2Config.tiny}[name]()
    model = GPT2(cfg)
    if device is not None:
        model = model.to(device)
    return model.to(dtype)
