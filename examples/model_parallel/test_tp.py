"""Tensor / tensor+sequence parallel block vs the serial block (reference:
examples/model_parallel/test_attn.py, test_tpmlp.py, test_transformer.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from _common import init, log, tdp
from torchdistpackage_b200.parallel import Block, ParallelBlock
from torchdistpackage_b200.parallel.tensor_parallel.transformer import allreduce_sequence_parallel_grads

rank, world, dev = init(__doc__)
dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
tol = 4e-2 if dtype == torch.bfloat16 else 1e-4
tdp.fix_rand(0)
dim, heads, B, N = 1024, 8, 4 * world, 128
serial = Block(dim, num_heads=heads).to(dev)
with torch.no_grad():
    for p in serial.parameters():
        if p.dim() == 2: p.mul_(0.08).sub_(0.04)
serial = serial.to(dtype)
x = torch.randn(B, N, dim, device=dev).to(dtype)
ys = serial(x)
for sp in (False, True):
    par = ParallelBlock(dim, num_heads=heads, sequence_parallel=sp).to(dev).to(dtype)
    par.init_from_full(serial)
    yp = par(x)
    ref = ys if not sp else ys.chunk(world)[rank]
    err = ((yp.float() - ref.float()).abs().max() / ref.float().abs().max()).item()
    assert err < tol, (sp, err)
    log(rank, f"sequence_parallel={sp}: rel err {err:.2e} OK")
