"""BASELINE.json config #3: 4-layer transformer, tensor parallel = world + sequence parallel,
h=4096, 32 heads, [B=4, N=2048] bf16 -- fwd + bwd + AdamW step, tokens/s (device-timed, max over
ranks).  `--impl ours|ours_nccl|reference` (reference = unmodified baseline/_ref TP layers).

    torchrun --nproc-per-node 8 scripts/bench_tp.py --impl ours
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="ours", choices=["ours", "ours_nccl", "reference"])
ap.add_argument("--dim", type=int, default=4096)
ap.add_argument("--heads", type=int, default=32)
ap.add_argument("--depth", type=int, default=4)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--seq", type=int, default=2048)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

if args.impl == "reference":
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import torchdistpackage as pkg
    from torchdistpackage.parallel.tensor_parallel.transformer import Transformer
    from torchdistpackage.parallel.tensor_parallel import tp_utils
    try:
        pkg.setup_distributed("nccl")
    except UnboundLocalError:
        pass            # reference defect under torchrun (process group is up)
else:
    import torchdistpackage_b200 as pkg
    from torchdistpackage_b200.parallel.tensor_parallel.transformer import (
        Transformer, allreduce_sequence_parallel_grads)
    from torchdistpackage_b200.parallel.tensor_parallel import tp_utils, tp_fused
    pkg.setup_distributed("nccl")
    pkg.tpc.verbose = False
    tp_fused.set_enabled(args.impl == "ours")

rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", torch.cuda.current_device())
# (the reference's hybrid-group pass needs a "data" entry: process_topo.py:114)
pkg.tpc.setup_process_groups([("tensor", world), ("data", 1)])
tp_utils.set_tp_group(pkg.tpc.get_group("tensor"))

torch.manual_seed(0)
model = Transformer(args.dim, mlp_ratio=4, num_heads=args.heads, depth=args.depth,
                    tensor_parallel=True, sequence_parallel=True)
with torch.no_grad():           # same sane init in every arm (the reference uses torch.rand)
    g = torch.Generator().manual_seed(1)
    for p in model.parameters():
        if p.dim() == 2:
            p.copy_(torch.randn(p.shape, generator=g) * 0.02)
model = model.to(dev).to(torch.bfloat16)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
if args.batch % world:      # sequence parallelism shards dim 0 of [B, N, h]
    args.batch = -(-args.batch // world) * world
x = torch.randn(args.batch, args.seq, args.dim, device=dev).to(torch.bfloat16)
dist.broadcast(x, 0)


def step():
    opt.zero_grad(set_to_none=True)
    out = model(x)
    loss = out.float().pow(2).mean()
    loss.backward()
    if args.impl != "reference":
        allreduce_sequence_parallel_grads(model)
    opt.step()
    return loss


for _ in range(args.warmup):
    step()
torch.cuda.synchronize(); dist.barrier()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(args.steps):
    loss = step()
e.record(); torch.cuda.synchronize()
t = torch.tensor([s.elapsed_time(e) / args.steps], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    tokens = args.batch * args.seq
    print(json.dumps({"config": "4-layer transformer TP=%d + SP (h=%d, %d heads, B=%d, N=%d)" % (
        world, args.dim, args.heads, args.batch, args.seq), "impl": args.impl, "n_gpus": world,
        "ms_per_step": t.item(), "tokens_per_s": tokens / (t.item() / 1e3), "dtype": "bf16",
        "loss": float(loss.item())}), flush=True)
dist.barrier()
dist.destroy_process_group()
