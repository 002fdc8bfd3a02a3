"""TpAttention vs the serial Attention: forward and input/weight gradients
(reference example: examples/model_parallel/test_attn.py).

    torchrun --nproc-per-node 2 examples/model_parallel/test_attn.py [--cpu]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from _common import init, log, tdp
from torchdistpackage_b200.parallel import Attention, TpAttention

rank, world, dev = init(__doc__)
dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
tol = 4e-2 if dtype == torch.bfloat16 else 1e-4
tdp.fix_rand(0)
dim, heads, B, N = 1024, 8, 2, 128
serial = Attention(dim, num_heads=heads).to(dev)
with torch.no_grad():
    for p in serial.parameters():
        if p.dim() == 2:
            p.mul_(0.08).sub_(0.04)
serial = serial.to(dtype)
par = TpAttention(dim, num_heads=heads).to(dev).to(dtype)
par.qkv.init_weight_from_full_attn(serial.qkv.weight, serial.qkv.bias)
par.proj.init_weight_from_full(serial.proj.weight, serial.proj.bias)

x = torch.randn(B, N, dim, device=dev).to(dtype)
xs, xp = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
ys, yp = serial(xs), par(xp)
gy = torch.randn_like(ys)
ys.backward(gy)
yp.backward(gy)


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


e_f, e_b = rel(yp, ys), rel(xp.grad, xs.grad)
d = dim // world
e_w = rel(par.proj.linear.weight.grad, serial.proj.weight.grad[rank * d:(rank + 1) * d])
assert max(e_f, e_b, e_w) < tol, (e_f, e_b, e_w)
log(rank, f"TpAttention: fwd {e_f:.2e}  dx {e_b:.2e}  dW_proj {e_w:.2e}  OK")
