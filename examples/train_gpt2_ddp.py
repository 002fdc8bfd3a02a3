"""End-to-end training loop with the pieces a production run uses together:

    NaiveDDP (symmetric-memory buckets, NVLS all-reduce overlapped with backward)
  + BucketAdamW (one fused kernel per bucket)   + GraphedStep (whole step as one CUDA graph, on GPU)
  + StepWatchdog (hang detection)  + MetricsLogger (JSONL)  + AsyncCheckpointWriter / resume.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_gpt2_ddp.py --model small --steps 200
    torchrun --nproc-per-node 2 examples/train_gpt2_ddp.py --cpu --model tiny --steps 6 --ckpt-every 3
    ... --resume            continues from the newest checkpoint in --out
"""
import argparse
import glob
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import torchdistpackage_b200 as tdp
from torchdistpackage_b200.dist.model_parallel_ckpt import AsyncCheckpointWriter, load_mp_checkpoint
from torchdistpackage_b200.models.gpt2 import build_gpt2
from torchdistpackage_b200.ops.fused import BucketAdamW
from torchdistpackage_b200.tools import MetricsLogger, StepWatchdog


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--model", default="tiny", choices=["tiny", "small", "medium"])
    ap.add_argument("--micro-batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--ckpt-every", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/train_gpt2_ddp")
    ap.add_argument("--resume", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args(argv)

    on_gpu = torch.cuda.is_available() and not args.cpu
    rank, world, _, _ = tdp.setup_distributed("nccl" if on_gpu else "gloo")
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    tdp.tpc.verbose = False
    tdp.tpc.setup_process_groups([("data", world)])
    tdp.fix_rand(0, deterministic_cudnn=False)

    model = build_gpt2(args.model, device=dev, dtype=torch.bfloat16 if on_gpu else torch.float32)
    ddp = tdp.NaiveDDP(model, sync=False, gradient_as_bucket_view=True, process_group=tdp.tpc.get_group("data"))
    opt = BucketAdamW(ddp, lr=args.lr, betas=(0.9, 0.95), weight_decay=0.1)

    start = 0
    if args.resume:
        found = sorted(glob.glob(args.out + "/step*.pth"))
        if found:
            prefix = found[-1][:-len(".pth")]
            state = load_mp_checkpoint(prefix)
            model.load_state_dict(state["model"])            # parameters are views of the flat buffers
            opt.load_state_dict(state["optimizer"])          # moments, step, fp32 masters (-> parameters)
            start = int(state["step"])
            if rank == 0:
                print(f"resumed from {prefix} at step {start}", flush=True)

    def train_step(tokens, targets):
        opt.zero_grad()
        loss = ddp(tokens, targets)
        loss.backward()
        ddp.reduce_gradients()
        opt.step()
        return loss

    step_fn = train_step
    if on_gpu and not args.no_graph:
        from torchdistpackage_b200.ops.graph import GraphedStep
        ex = torch.zeros(args.micro_batch, model.cfg.seq_len, dtype=torch.long, device=dev)
        # the capture's warm-up iterations are real optimizer steps on an all-zero batch:
        # `preserve` snapshots parameters + optimizer state (incl. the device step counter) around
        # them, so a graph built after a resume continues exactly where the checkpoint left off
        step_fn = GraphedStep(train_step, (ex, ex.clone()), warmup=2, preserve=opt.state_tensors())

    writer = AsyncCheckpointWriter()
    metrics = MetricsLogger(os.path.join(args.out, "metrics.jsonl"))
    tokens_per_step = world * args.micro_batch * model.cfg.seq_len
    with StepWatchdog(timeout_s=600, rank=rank, world=world) as wd:
        for step in range(start, args.steps):
            gen = torch.Generator().manual_seed(1000 * step + rank)     # data is a function of the step
            tok = torch.randint(0, model.cfg.vocab_size, (args.micro_batch, model.cfg.seq_len + 1), generator=gen)
            tok = tok.pin_memory().to(dev, non_blocking=True) if on_gpu else tok
            t0 = time.perf_counter()
            loss = step_fn(tok[:, :-1].contiguous(), tok[:, 1:].contiguous())
            lv = float(loss.detach())                                   # D2H read of the step's result
            wd.tick(step)
            metrics.log(step, loss=lv, tokens_per_s=tokens_per_step / (time.perf_counter() - t0))
            if rank == 0:
                print(f"step {step} loss {lv:.4f}", flush=True)
            if args.ckpt_every and (step + 1) % args.ckpt_every == 0:
                writer.save(os.path.join(args.out, f"step{step + 1:06d}"),
                            {"model": model.state_dict(), "optimizer": opt.state_dict(), "step": step + 1})
    writer.wait()
    metrics.close()
    if dist.is_initialized():
        dist.barrier()
    return model


if __name__ == "__main__":
    main()
