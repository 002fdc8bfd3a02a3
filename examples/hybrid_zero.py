"""Hybrid ZeRO: shard optimizer state inside a node (NVSwitch), plain data parallel across nodes
(reference: dist/node_group.py docstring, Intro.md:69-79).  On a single 8-GPU box
setup_node_groups() returns None and the data group is used for both roles."""
import torch, torch.nn as nn
from _common import init, log, tdp

rank, world, dev = init(__doc__)
tdp.tpc.setup_process_groups([("data", world)])
node_group = tdp.setup_node_groups(num_per_node=8)
model = nn.Sequential(nn.Linear(256, 1024), nn.GELU(), nn.Linear(1024, 256)).to(dev)
if node_group is not None:
    # multi-node: DDP averages across nodes (same local rank), ZeRO shards inside the node
    inter = tdp.tpc.get_group("data")
    model = tdp.NaiveDDP(model, process_group=inter, gradient_as_bucket_view=True)
zopt = tdp.Bf16ZeroOptimizer(torch.optim.AdamW(model.parameters(), lr=1e-3),
                             dp_group=node_group or tdp.tpc.get_group("data"), overlap_comm=True)
for it in range(3):
    x = torch.randn(32, 256, device=dev)
    zopt.zero_grad(); loss = model(x).pow(2).mean(); loss.backward()
    if node_group is not None:
        model.reduce_gradients()
    zopt.step()
    log(rank, f"step {it} loss {loss.item():.4f}")
