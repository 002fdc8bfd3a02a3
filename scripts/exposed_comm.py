"""Exposed communication per step of the flagship data-parallel run (BASELINE metric: "tokens/s
... plus exposed communication ms/step"): the same CUDA-graphed GPT-2 small step is timed with the
bucket all-reduces enabled and with the collective skipped (everything else identical); the
difference is what communication costs the step after overlap.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/exposed_comm.py [--steps 20]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchdistpackage_b200 as tdp  # noqa: E402
from torchdistpackage_b200.models.gpt2 import build_gpt2  # noqa: E402
from torchdistpackage_b200.ops.fused import BucketAdamW  # noqa: E402
from torchdistpackage_b200.ops.graph import GraphedStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="small")
ap.add_argument("--micro-batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=5)
args = ap.parse_args()

rank, world, _, _ = tdp.setup_distributed("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
tdp.tpc.verbose = False
tdp.tpc.setup_process_groups([("data", world)])
pg = tdp.tpc.get_group("data")


def timed(skip_comm: bool) -> float:
    tdp.fix_rand(0, deterministic_cudnn=False)
    model = build_gpt2(args.model, device=dev)
    ddp = tdp.NaiveDDP(model, sync=False, gradient_as_bucket_view=True, bucket_cap_mb=25, process_group=pg)
    if skip_comm:
        red = ddp.reducer
        red._reduce_bucket = lambda bucket: setattr(bucket, "reduced", True)   # no collective
    opt = BucketAdamW(ddp, lr=3e-4, betas=(0.9, 0.95), weight_decay=0.1)
    gen = torch.Generator().manual_seed(1 + rank)
    toks = [torch.randint(0, model.cfg.vocab_size, (args.micro_batch, model.cfg.seq_len + 1), generator=gen).to(dev)
            for _ in range(4)]

    def eager(tokens, targets):
        opt.zero_grad()
        loss = ddp(tokens, targets)
        loss.backward()
        ddp.reduce_gradients()
        opt.step()
        return loss

    g = GraphedStep(eager, (toks[0][:, :-1].contiguous(), toks[0][:, 1:].contiguous()), warmup=2)
    for i in range(args.warmup):
        g(toks[i % 4][:, :-1].contiguous(), toks[i % 4][:, 1:].contiguous())
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for i in range(args.steps):
        g(toks[i % 4][:, :-1].contiguous(), toks[i % 4][:, 1:].contiguous())
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / args.steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del g, opt, ddp, model
    torch.cuda.empty_cache()
    return t.item()


t_comm = timed(False)
t_nocomm = timed(True)
if rank == 0:
    out = {"metric": "exposed communication per step (DDP bucket all-reduce, after overlap)", "n_gpus": world,
           "model": "gpt2-" + args.model, "ms_per_step": t_comm, "ms_per_step_without_collective": t_nocomm,
           "exposed_comm_ms": t_comm - t_nocomm, "steps": args.steps}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/exposed_comm_w{world}.json", "w"), indent=1)
    print(json.dumps(out), flush=True)
dist.barrier()
