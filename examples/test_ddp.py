"""NaiveDDP must track torch DistributedDataParallel step for step (reference: examples/test_ddp.py),
here with DIFFERENT data on every rank so the reduction is really exercised."""
import copy
import torch, torch.nn as nn
from torch.nn.parallel import DistributedDataParallel as TorchDDP
from _common import init, log, tdp

rank, world, dev = init(__doc__)
tdp.fix_rand(0)
model = nn.Sequential(nn.Linear(10, 10), nn.ReLU(), nn.Linear(10, 1)).to(dev)
ref = TorchDDP(copy.deepcopy(model), device_ids=[dev.index] if dev.type == "cuda" else None)
ddp = tdp.NaiveDDP(model, sync=False, gradient_as_bucket_view=True)
opt = torch.optim.Adam(ddp.parameters(), lr=1e-2)
ropt = torch.optim.Adam(ref.parameters(), lr=1e-2)
for it in range(10):
    torch.manual_seed(1000 * it + rank)
    x = torch.randn(3, 10, device=dev)
    ddp(x).sum().backward(); ddp.reduce_gradients()
    ref(x).sum().backward()
    for p, q in zip(ddp.module.parameters(), ref.module.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-5), it
    opt.step(); ropt.step(); opt.zero_grad(); ropt.zero_grad()
    for p, q in zip(ddp.module.parameters(), ref.module.parameters()):
        assert torch.allclose(p, q, atol=1e-5)
log(rank, "NaiveDDP == torch DDP for 10 Adam steps: OK")
