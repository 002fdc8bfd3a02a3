"""One launch of every bandwidth-bound kernel of the GPT-2-small step at its real shape (for ncu):
LayerNorm fwd / bwd, softmax-cross-entropy fwd+bwd, AdamW (one 25 MiB bucket), bias-gradient column
sums, packed-gradient row copy, multi-tensor sumsq / scale.   python scripts/misc_kernels_one.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdistpackage_b200.ops._loader import native  # noqa: E402
from torchdistpackage_b200.ops import fused as Fo  # noqa: E402

C = native(required=True)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
T, D, V = 16384, 768, 50304
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

x = torch.randn(T, D, device=dev).to(torch.bfloat16).requires_grad_(True)
w = torch.ones(D, device=dev, dtype=torch.bfloat16, requires_grad=True)
b = torch.zeros(D, device=dev, dtype=torch.bfloat16, requires_grad=True)
flush.zero_()
y = Fo.layer_norm(x, w, b, 1e-5)                       # layernorm_fwd
flush.zero_()
y.backward(torch.randn_like(y))                       # layernorm_bwd + partial reduce

logits = torch.randn(T, V, device=dev).to(torch.bfloat16)
tgt = torch.randint(0, V, (T,), device=dev)
loss = torch.zeros(T, device=dev)
flush.zero_()
C.cross_entropy_fwd_bwd(logits, tgt, loss, 1.0 / T, -100)   # softmax-CE forward + d(logits) in place

n = 25 * (1 << 20) // 2                                # one DDP bucket worth of bf16 parameters
p = torch.randn(n, device=dev).to(torch.bfloat16)
g = torch.randn(n, device=dev).to(torch.bfloat16)
m = p.float().clone(); ea = torch.zeros(n, device=dev); es = torch.zeros(n, device=dev)
flush.zero_()
C.adamw(p, m, g, ea, es, 3e-4, 0.9, 0.95, 1e-8, 0.1, 1, True, 1.0, None, None)

dy = torch.randn(T, 4 * D, device=dev).to(torch.bfloat16)
out = torch.empty(4 * D, device=dev, dtype=torch.bfloat16)
flush.zero_()
C.colsum(dy, out)                                      # bias gradient of fc1

src = torch.randn(16, 12, 1024, 64, device=dev).to(torch.bfloat16)
dst = torch.empty(16, 1024, 3, 12, 64, device=dev, dtype=torch.bfloat16)
flush.zero_()
C.permute_rows_copy(dst[:, :, 0].transpose(1, 2), src)  # dq -> packed dqkv window

grads = [torch.randn(s, device=dev).to(torch.bfloat16) for s in (D * 3 * D, D * D, 4 * D * D, 4 * D * D, D, 3 * D)]
flush.zero_()
tot = Fo.multi_sumsq(grads)
Fo.multi_scale_(grads, 1.0, torch.ones(1, device=dev))
torch.cuda.synchronize()
print("done", float(loss.mean()), float(tot))
