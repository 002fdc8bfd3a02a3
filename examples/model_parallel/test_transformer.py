"""TP + SP Transformer (stack of ParallelBlocks) vs the serial Transformer
(reference example: examples/model_parallel/test_transformer.py -- without the stray debugger
breakpoint, and with the sequence-parallel LayerNorm gradients all-reduced).

    torchrun --nproc-per-node 2 examples/model_parallel/test_transformer.py [--cpu]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from _common import init, log, tdp
from torchdistpackage_b200.parallel import Transformer
from torchdistpackage_b200.parallel.tensor_parallel.transformer import allreduce_sequence_parallel_grads

rank, world, dev = init(__doc__)
dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
tol = 6e-2 if dtype == torch.bfloat16 else 1e-4
dim, heads, depth, B, N = 512, 8, 2, 2 * world, 64


def build(tp, sp):
    tdp.fix_rand(0)
    m = Transformer(dim, num_heads=heads, depth=depth, tensor_parallel=tp, sequence_parallel=sp)
    return m.to(dev)


serial = build(False, False)
with torch.no_grad():
    for p in serial.parameters():
        if p.dim() == 2:
            p.mul_(0.08).sub_(0.04)
par = build(True, True)
for bs, bp in zip(serial.blocks, par.blocks):
    bp.init_from_full(bs)
serial, par = serial.to(dtype), par.to(dtype)

x = torch.randn(B, N, dim, device=dev).to(dtype)
ys = serial(x)
yp = par(x)                 # the model shards the sequence itself and gathers it back
err = ((yp.float() - ys.float()).abs().max() / ys.float().abs().max()).item()
assert err < tol, err
yp.float().mean().backward()
allreduce_sequence_parallel_grads(par)
log(rank, f"TP+SP transformer depth={depth}: rel err {err:.2e}  OK")
