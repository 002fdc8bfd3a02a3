"""Reference arm of bench.py: the UNMODIFIED KimmiShi/TorchDistPackage (installed under
``baseline/_ref`` with ``pip install --target``) driving a plain-PyTorch GPT-2 of the same
architecture / shapes / dtype / optimizer as the B200-native arm.

Nothing from ``torchdistpackage_b200`` is imported here: the model is ordinary ``torch.nn`` code
(nn.Linear, nn.LayerNorm, F.scaled_dot_product_attention, F.gelu(tanh), F.cross_entropy -- i.e.
cuBLAS / ATen / flash-SDPA kernels), data parallelism is the reference's own
``torchdistpackage.NaiveDDP`` (bucketed NCCL all-reduce on a side stream, stock code path:
``gradient_as_bucket_view=True`` + ``reduce_gradients()``), process groups come from the
reference's ``setup_distributed`` + ``tpc.setup_process_groups``, the optimizer is
``torch.optim.AdamW(fused=True)``.

Harness caveats: ``setup_distributed`` raises
``UnboundLocalError`` under torchrun *after* the process group is up -- caught here;
bucket-view mode needs ``zero_grad(set_to_none=False)``.
"""
from __future__ import annotations

import math
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
if _REF not in sys.path:
    sys.path.insert(0, _REF)

import torchdistpackage as ref_pkg  # noqa: E402  (the reference, unmodified)


class _Cfg:
    def __init__(self, name):
        self.vocab_size, self.seq_len, self.mlp_ratio = 50304, 1024, 4
        if name == "small":
            self.n_layer, self.n_head, self.d_model = 12, 12, 768
        elif name == "medium":
            self.n_layer, self.n_head, self.d_model = 24, 16, 1024
        else:
            self.vocab_size, self.seq_len = 512, 128
            self.n_layer, self.n_head, self.d_model = 2, 4, 128

    def flops_per_token(self) -> float:
        d, L = self.d_model, self.n_layer
        matmul_params = L * (4 * d * d + 2 * self.mlp_ratio * d * d) + self.vocab_size * d
        attn = L * 2 * 2 * self.seq_len * d / 2
        return 6.0 * matmul_params + 3.0 * attn


class _Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        d = c.d_model
        self.n_head = c.n_head
        self.ln_1, self.ln_2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.qkv, self.proj = nn.Linear(d, 3 * d), nn.Linear(d, d)
        self.fc1, self.fc2 = nn.Linear(d, c.mlp_ratio * d), nn.Linear(c.mlp_ratio * d, d)
        for m in (self.qkv, self.fc1):
            nn.init.normal_(m.weight, std=0.02)
        for m in (self.proj, self.fc2):
            nn.init.normal_(m.weight, std=0.02 / math.sqrt(2 * c.n_layer))
        for m in (self.qkv, self.proj, self.fc1, self.fc2):
            nn.init.zeros_(m.bias)

    def forward(self, x):
        B, T, D = x.shape
        q, k, v = self.qkv(self.ln_1(x)).view(B, T, 3, self.n_head, D // self.n_head).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, T, D)
        x = x + self.proj(o)
        return x + self.fc2(F.gelu(self.fc1(self.ln_2(x)), approximate="tanh"))


class TorchGPT2(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.cfg = c
        self.wte = nn.Embedding(c.vocab_size, c.d_model)
        self.wpe = nn.Embedding(c.seq_len, c.d_model)
        self.blocks = nn.ModuleList([_Block(c) for _ in range(c.n_layer)])
        self.ln_f = nn.LayerNorm(c.d_model)
        nn.init.normal_(self.wte.weight, std=0.02)
        nn.init.normal_(self.wpe.weight, std=0.02)

    def forward(self, idx, targets):
        x = self.wte(idx) + self.wpe(torch.arange(idx.shape[1], device=idx.device))
        for b in self.blocks:
            x = b(x)
        logits = F.linear(self.ln_f(x), self.wte.weight)
        return F.cross_entropy(logits.view(-1, logits.shape[-1]), targets.reshape(-1))


def init_distributed():
    """Bring the process group up through the reference's own ``setup_distributed``."""
    try:
        ref_pkg.setup_distributed(backend="nccl")
    except UnboundLocalError:
        # reference defect (dist/launch_from_slurm.py:62): `addr` is unbound under torchrun; the
        # process group has been initialised by then.
        pass


def build_reference(args, device, world):
    import torch.distributed as dist
    if world == 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        os.environ["RANK"], os.environ["WORLD_SIZE"] = "0", "1"
        init_distributed()
    torch.manual_seed(2222)
    cfg = _Cfg(args.model)
    model = TorchGPT2(cfg).to(device).to(torch.bfloat16)
    ref_pkg.tpc.setup_process_groups([("data", world)])
    group = ref_pkg.tpc.get_group("data")
    ddp = ref_pkg.NaiveDDP(model, sync=False, bucket_cap_mb=25, gradient_as_bucket_view=True,
                           process_group=group, dp_rank0=0)
    opt = torch.optim.AdamW(model.parameters(), lr=3e-4, betas=(0.9, 0.95), eps=1e-8,
                            weight_decay=0.1, fused=True)

    def step(tokens, targets):
        opt.zero_grad(set_to_none=False)
        loss = ddp(tokens, targets)
        loss.backward()
        ddp.reduce_gradients()
        opt.step()
        return loss

    return step, (lambda: 0), cfg
