"""FSDP2 (``torch.distributed.fsdp.fully_shard``) + CPU parameter offload memory experiment
(reference example: examples/fsdp2_offload_test.py, which does this with a HuggingFace
Qwen2.5-VL checkpoint; there is no network here, so the subject is this package's GPT-2).

Not a package feature -- it is the torch-native baseline to compare Bf16ZeroOptimizer /
hybrid ZeRO against: peak device memory and step time with (a) plain FSDP2, (b) FSDP2 with
``CPUOffloadPolicy`` (parameters, gradients and optimizer state live in pinned host memory and
stream through the GPU layer by layer), (c) this package's ZeRO optimizer.

    torchrun --nproc-per-node 2 examples/fsdp2_offload_test.py [--model medium] [--steps 5]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist

import torchdistpackage_b200 as tdp
from torchdistpackage_b200.models.gpt2 import build_gpt2


def peak_gb():
    return torch.cuda.max_memory_allocated() / 2 ** 30


def run(name, model, opt, tokens, steps, step_fn=None):
    torch.cuda.reset_peak_memory_stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = model(tokens[:, :-1], tokens[:, 1:])
        loss.backward()
        (step_fn or opt.step)()
        opt.zero_grad()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps * 1e3
    if dist.get_rank() == 0:
        print(f"{name:28s} peak {peak_gb():6.2f} GiB   {dt:8.1f} ms/step   loss {float(loss):.3f}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="small", choices=["tiny", "small", "medium"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=1024)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        print("needs CUDA devices (FSDP2 offload streams parameters through the GPU)")
        return
    # a CPU (gloo) group next to NCCL: the offloaded optimizer step reduces on the host
    tdp.setup_distributed("cpu:gloo,cuda:nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    from torch.distributed.fsdp import CPUOffloadPolicy, MixedPrecisionPolicy, fully_shard
    tokens = torch.randint(0, 50257, (args.batch, args.seq + 1), device=dev)
    mp = MixedPrecisionPolicy(param_dtype=torch.bfloat16, reduce_dtype=torch.float32)

    for offload in (False, True):
        tdp.fix_rand(0)
        model = build_gpt2(args.model, device="cpu" if offload else dev, dtype=torch.float32)
        kw = dict(mp_policy=mp)
        if offload:
            kw["offload_policy"] = CPUOffloadPolicy(pin_memory=True)
        for blk in model.blocks:
            fully_shard(blk, **kw)
        fully_shard(model, **kw)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        run(f"FSDP2{' + CPU offload' if offload else ''}", model, opt, tokens, args.steps)
        del model, opt
        torch.cuda.empty_cache()

    tdp.fix_rand(0)
    model = build_gpt2(args.model, device=dev, dtype=torch.bfloat16)
    opt = tdp.Bf16ZeroOptimizer(torch.optim.AdamW(model.parameters(), lr=1e-4), overlap_comm=True)
    run("Bf16ZeroOptimizer (this repo)", model, opt, tokens, args.steps)


if __name__ == "__main__":
    main()
