"""Bf16ZeroOptimizer (ZeRO shards + reduce-scatter / all-gather) vs torch DDP + Adam
(reference: examples/test_zero_optim.py)."""
import copy
import torch, torch.nn as nn
from torch.nn.parallel import DistributedDataParallel as TorchDDP
from _common import init, log, tdp

rank, world, dev = init(__doc__)
tdp.fix_rand(0)
model = nn.Sequential(nn.Linear(64, 256), nn.GELU(), nn.Linear(256, 8)).to(dev)
ref = TorchDDP(copy.deepcopy(model), device_ids=[dev.index] if dev.type == "cuda" else None)
zopt = tdp.Bf16ZeroOptimizer(torch.optim.Adam(model.parameters(), lr=1e-3), overlap_comm=True)
ropt = torch.optim.Adam(ref.parameters(), lr=1e-3)
for it in range(10):
    torch.manual_seed(77 * it + rank)
    x = torch.randn(16, 64, device=dev) + rank
    zopt.zero_grad(); model(x).pow(2).mean().backward(); zopt.step()
    ropt.zero_grad(); ref(x).pow(2).mean().backward(); ropt.step()
    for p, q in zip(model.parameters(), ref.module.parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), it
log(rank, "ZeRO == DDP+Adam for 10 steps: OK")
