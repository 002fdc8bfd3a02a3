"""ShardedEMA vs a full-replica EMA (reference: examples/test_shard_ema.py)."""
import torch, torch.nn as nn
from _common import init, log, tdp

rank, world, dev = init(__doc__)
tdp.test_comm(verbose=rank == 0)
tdp.fix_rand(0)
model = nn.Sequential(*[nn.Linear(128, 128) for _ in range(8)]).to(dev)
ema = tdp.ShardedEMA(model)
full = {n: p.detach().clone() for n, p in model.named_parameters()}
for it in range(100):
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    ema.update(model, decay=0.999)
    for n, p in model.named_parameters():
        full[n].mul_(0.999).add_(p.detach(), alpha=0.001)
assert ema.verify_with_gt(full, rtol=1e-5, atol=1e-6)
sd = ema.state_dict_cpu()
if rank == 0:
    assert all(torch.allclose(sd[n], full[n].cpu(), atol=1e-6) for n in full)
log(rank, "ShardedEMA == full EMA after 100 updates: OK")
