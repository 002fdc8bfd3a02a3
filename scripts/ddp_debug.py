"""Root-cause helper for the 8-GPU `ddp_grad_vs_nccl_avg` failure seen in round 1
(`profiles/engines_check_w8_r1.log`).  torchrun --nproc-per-node N scripts/ddp_debug.py

Separates the suspects:
  A. the in-place NVLS all-reduce on DDP-sized, oddly sized payloads (no autograd, no overlap),
     repeated back to back on one stream with a zero-fill in between (the late-store hypothesis);
  B. NaiveDDP with the reduction forced to the end of backward (sync=True) vs overlapped;
  C. which parameters / buckets differ, on which ranks, and by how much.
"""
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchdistpackage_b200 as tdp  # noqa: E402
from torchdistpackage_b200.models.gpt2 import build_gpt2  # noqa: E402
from torchdistpackage_b200.ops.symm import get_symm_group  # noqa: E402

rank, world, _, _ = tdp.setup_distributed("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
tdp.tpc.verbose = False
tdp.tpc.setup_process_groups([("data", world)])
dp = tdp.tpc.get_group("data")
res = {"world": world}


def log(*a):
    if rank == 0:
        print(*a, flush=True)


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


# ---------------------------------------------------------------- A: raw all-reduce stress
sg = get_symm_group(dp)
buf = sg.alloc(8 << 20)
worst_a = 0.0
for n in (8, 264, 66432, 131072, 131968, 1 << 20):
    for rep in range(20):
        x = torch.randn(n, device=dev).to(torch.bfloat16)
        v = buf.view(0, (n,), torch.bfloat16)
        v.zero_()                                   # local write right after the previous reduction
        v.copy_(x)
        buf.all_reduce_(0, n, torch.bfloat16, 1.0 / world)
        ref = x.float().clone()
        dist.all_reduce(ref)
        ref /= world
        torch.cuda.synchronize()
        r = rel(v, ref)
        worst_a = max(worst_a, r)
        if r > 2e-2:
            bad = ((v.float() - ref).abs() > 2e-2 * ref.abs().max()).nonzero().flatten()
            print(f"[rank {rank}] raw all-reduce n={n} rep={rep}: rel {r:.3e}, {bad.numel()} bad elements, "
                  f"first {bad[:4].tolist()} last {bad[-4:].tolist()}", flush=True)
res["raw_allreduce_worst"] = worst_a
log("A raw in-place all-reduce, worst rel:", worst_a)

# ---------------------------------------------------------------- B / C: NaiveDDP variants
for sync in (True, False):
    for cap in (0.25, 25.0):
        tdp.fix_rand(0, deterministic_cudnn=False)
        model = build_gpt2("tiny", device=dev)
        ref = copy.deepcopy(model)
        ddp = tdp.NaiveDDP(model, sync=sync, gradient_as_bucket_view=True, process_group=dp, bucket_cap_mb=cap)
        torch.manual_seed(100 + rank)
        tok = torch.randint(0, model.cfg.vocab_size, (4, model.cfg.seq_len + 1), device=dev)
        for it in range(3):
            ddp.zero_grad()
            ddp(tok[:, :-1], tok[:, 1:]).backward()
            ddp.reduce_gradients()
        ref(tok[:, :-1], tok[:, 1:]).backward()
        bad = []
        for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            g = q.grad.float().clone()
            dist.all_reduce(g)
            g /= world
            r = rel(p.grad, g)
            if r > 3e-2:
                b = ddp.reducer.param_bucket[n]
                bad.append((n, round(r, 3), b.index, float((p.grad.float() / g.clamp_min(1e-12)).median())))
        worst = torch.tensor([len(bad)], device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        print(f"[rank {rank}] sync={sync} cap={cap}: {len(bad)} mismatching params {bad[:6]}", flush=True)
        res[f"ddp_sync{int(sync)}_cap{cap}"] = int(worst.item())
        del ddp, model, ref
        torch.cuda.synchronize()
        dist.barrier()

if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/ddp_debug_w{world}.json", "w"), indent=1)
    print("DONE", res, flush=True)
dist.barrier()
