"""1F1B pipeline x data parallel smoke train (reference: examples/model_parallel/test_pipeline.py):
a Sequential is split over `pp` stages, each stage is wrapped in NaiveDDP with
num_grad_acc_iter = #micro-batches so gradients reduce once per mini-batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from _common import init, log, tdp
from torchdistpackage_b200.parallel import forward_backward, partition_uniform

rank, world, dev = init(__doc__)
pp = 2 if world % 2 == 0 else 1
tdp.tpc.setup_process_groups([("data", world // pp), ("pipe", pp)])
tdp.fix_rand(0)
layers = [nn.Linear(10, 10) for _ in range(6)]
stage = nn.Sequential(*partition_uniform(layers)).to(dev)
n_micro = 4
ddp = tdp.NaiveDDP(stage, process_group=tdp.tpc.get_group("data"), gradient_as_bucket_view=True,
                   num_grad_acc_iter=n_micro, dp_rank0=tdp.tpc.get_ranks_in_group("data")[0])
opt = torch.optim.SGD(stage.parameters(), lr=0.05)
first, last = tdp.tpc.is_first_in_pipeline_group(), tdp.tpc.is_last_in_pipeline_group()

def fwd(inp):
    if last and pp > 1:
        act, tgt = inp
        return (ddp(act) - tgt).pow(2).mean()
    if pp == 1:
        return (ddp(inp[0]) - inp[1]).pow(2).mean()
    return ddp(inp)

for epoch in range(2):
    for it in range(5):
        x = torch.randn(512, 10, device=dev); y = torch.randn(512, 10, device=dev)
        inputs = ([x] if first else []) + ([y] if last else [])
        out = forward_backward(opt, fwd, None, inputs or None, num_microbatches=n_micro,
                               dtype=torch.float32)
        ddp.reduce_gradients(); opt.step()
        if last and tdp.tpc.get_dp_rank() == 0:
            print(f"epoch {epoch} it {it} loss(last micro-batch) {out.item():.4f}", flush=True)
log(rank, "pipeline example done")
