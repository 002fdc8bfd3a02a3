"""TpMlp vs the serial Mlp: forward and gradients, with and without sequence parallelism
(reference example: examples/model_parallel/test_tpmlp.py).

    torchrun --nproc-per-node 2 examples/model_parallel/test_tpmlp.py [--cpu]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from _common import init, log, tdp
from torchdistpackage_b200.parallel import Mlp, TpMlp

rank, world, dev = init(__doc__)
dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
tol = 4e-2 if dtype == torch.bfloat16 else 1e-4
tdp.fix_rand(0)
dim, T = 1024, 512 * world
serial = Mlp(dim, hidden_features=4 * dim).to(dev)
with torch.no_grad():
    for p in serial.parameters():
        if p.dim() == 2:
            p.mul_(0.08).sub_(0.04)
serial = serial.to(dtype)


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


x = torch.randn(T, dim, device=dev).to(dtype)
gy = torch.randn(T, dim, device=dev).to(dtype)
xs = x.clone().requires_grad_(True)
ys = serial(xs)
ys.backward(gy)
for sp in (False, True):
    par = TpMlp(dim, hidden_features=4 * dim, sequence_parallel=sp).to(dev).to(dtype)
    par.fc1.init_weight_from_full(serial.fc1.weight, serial.fc1.bias)
    par.fc2.init_weight_from_full(serial.fc2.weight, serial.fc2.bias)
    xin = (x.chunk(world)[rank] if sp else x).clone().requires_grad_(True)
    yp = par(xin)
    yp.backward(gy.chunk(world)[rank] if sp else gy)
    ref_y = ys.chunk(world)[rank] if sp else ys
    ref_dx = xs.grad.chunk(world)[rank] if sp else xs.grad
    e_f, e_b = rel(yp, ref_y), rel(xin.grad, ref_dx)
    assert max(e_f, e_b) < tol, (sp, e_f, e_b)
    log(rank, f"TpMlp sequence_parallel={sp}: fwd {e_f:.2e}  dx {e_b:.2e}  OK")
