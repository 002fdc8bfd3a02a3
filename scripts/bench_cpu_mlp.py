"""BASELINE config #1: NaiveDDP on a 2-layer MLP, world_size 2, CPU / gloo (plumbing check).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/bench_cpu_mlp.py

The model and input are the reference's (examples/test_ddp.py:12-20,74-76: Linear(10,10) ->
Linear(10,1), input [3,10] fp32).  The reference itself cannot run this config: its NaiveDDP
creates a CUDA stream in the constructor and uses ReduceOp.AVG, which gloo does not have
(ddp/naive_ddp.py:53,76) -- ``--impl reference`` reports that.  Prints one JSON line (rank 0):
steps/s through the public API plus a gradient check against brute-force averaging."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--warmup", type=int, default=20)
args = ap.parse_args()

if args.impl == "reference":
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"impl": "reference", "config": "#1 NaiveDDP 2-layer MLP, world 2, CPU/gloo",
                          "unavailable": "the reference's NaiveDDP needs CUDA (torch.cuda.Stream() in its "
                                         "constructor, ReduceOp.AVG): naive_ddp.py:53,76"}))
    sys.exit(0)

import torchdistpackage_b200 as tdp  # noqa: E402

rank, world, _, _ = tdp.setup_distributed("gloo")
tdp.fix_rand(0)


class MLP(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(10, 10), nn.Linear(10, 1)

    def forward(self, x):
        return self.fc2(self.fc1(x))


model = MLP()
ddp = tdp.NaiveDDP(model, sync=False, gradient_as_bucket_view=True)
opt = torch.optim.Adam(ddp.parameters(), lr=1.5e-4)


def step(x):
    opt.zero_grad(set_to_none=False)
    loss = ddp(x).sum()
    loss.backward()
    ddp.reduce_gradients()
    opt.step()
    return loss


g = torch.Generator().manual_seed(100 + rank)
xs = [torch.rand(3, 10, generator=g) for _ in range(args.steps + args.warmup)]
for i in range(args.warmup):
    step(xs[i])
dist.barrier()
t0 = time.perf_counter()
for i in range(args.steps):
    loss = step(xs[args.warmup + i])
dist.barrier()
sec = time.perf_counter() - t0

# gradient proof: one more backward, compare with the average of every rank's local gradient
x = xs[-1]
opt.zero_grad(set_to_none=False)
ddp(x).sum().backward()
ddp.reduce_gradients()
got = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
ref = MLP()
ref.load_state_dict(model.state_dict())
ref(x).sum().backward()
local = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
dist.all_reduce(local)
err = float((got - local / world).abs().max())
t = torch.tensor([sec, err], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"impl": "ours", "config": "#1 NaiveDDP 2-layer MLP (10->10->1), input [3,10] fp32, "
                                                f"world {world}, CPU/gloo",
                      "metric": "steps/s (host-timed: a CPU run has no device clock), max over ranks",
                      "value": args.steps / float(t[0]), "ms_per_step": float(t[0]) / args.steps * 1e3,
                      "steps": args.steps, "warmup": args.warmup, "n_procs": world,
                      "grad_check_max_abs": float(t[1]), "final_loss": float(loss.detach())}))
dist.barrier()
tdp.shutdown_distributed()
