"""BASELINE.json config #4: MoE 8-expert transformer, EP=4 x moe-DP=2 on 8 GPUs (EP = world/2 on
smaller boxes): fwd + bwd + dense-grad all-reduce (data group) + expert-grad all-reduce (moe_dp
group) + AdamW, tokens/s device-timed (max over ranks).

  --impl ours       MoELayer with P2P dispatch/combine kernels, NaiveDDP on NVLS buckets, MoEDP
  --impl ours_a2a   same model, dispatch/combine through dist.all_to_all_single (NCCL)
  --impl reference  the reference provides only groups + gradient hooks: the MoE
                    layer is supplied by the harness (plain torch + all_to_all_single, the same
                    gate / capacity / layout) and is driven with the reference's
                    tpc.build_moe_groups + NaiveDDP; expert grads are reduced by a reference
                    NaiveDDP over the moe_dp group on an expert-only module (its MoEDP default path
                    performs no reduction -- naive_ddp.py:316-322).
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="ours", choices=["ours", "ours_a2a", "reference"])
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--dim", type=int, default=1024)
ap.add_argument("--heads", type=int, default=16)
ap.add_argument("--experts", type=int, default=8)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--seq", type=int, default=1024)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

if args.impl == "reference":
    os.environ["TDP_DISABLE_NATIVE"] = "1"   # reference arm: plain torch kernels (cuBLAS / ATen / NCCL)
import torchdistpackage_b200 as tdp          # model definition (shared by all arms)
from torchdistpackage_b200.models import MoETransformer, MoEConfig
from torchdistpackage_b200.moe import layer as moe_layer

if args.impl == "reference":
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import torchdistpackage as ref
    try:
        ref.setup_distributed("nccl")
    except UnboundLocalError:
        pass
    pkg = ref
else:
    tdp.setup_distributed("nccl")
    tdp.tpc.verbose = False
    pkg = tdp

rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", torch.cuda.current_device())
ep = 4 if world >= 8 else max(world // 2, 1)
pkg.tpc.setup_process_groups([("data", world)])
if world > 1:
    pkg.tpc.build_moe_groups(moe_ep_size=ep)
    ep_group, dp_group = pkg.tpc.get_group("moe_ep"), pkg.tpc.get_group("moe_dp")
    moe_dp_rank0 = pkg.tpc.get_ranks_in_group("moe_dp")[0]
else:
    ep_group = dp_group = None
    moe_dp_rank0 = 0

cfg = MoEConfig(n_layer=args.layers, n_head=args.heads, d_model=args.dim, seq_len=args.seq,
                num_experts=args.experts, top_k=2, capacity_factor=1.25)
torch.manual_seed(0)
model = MoETransformer(cfg, ep_group=ep_group).to(dev).to(torch.bfloat16)
if args.impl in ("ours_a2a", "reference"):
    # force the collective (NCCL all_to_all_single) dispatch/combine path
    orig = moe_layer._A2AContext.__init__
    def no_symm(self, group, hidden):
        orig(self, group, hidden)
        self.sym = None
    moe_layer._A2AContext.__init__ = no_symm

experts = model.expert_parameters()
model._ddp_params_and_buffers_to_ignore = list(experts.keys())
data_group = pkg.tpc.get_group("data")
if args.impl == "reference":
    ddp = ref.NaiveDDP(model, sync=False, gradient_as_bucket_view=True, process_group=data_group)

    class ExpertOnly(torch.nn.Module):
        def __init__(self, params):
            super().__init__()
            self.ps = torch.nn.ParameterList(list(params.values()))
    expert_ddp = ref.NaiveDDP(ExpertOnly(experts), sync=False, gradient_as_bucket_view=True,
                              process_group=dp_group, dp_rank0=moe_dp_rank0) if world > 1 else None
    def finish():
        ddp.reduce_gradients()
        if expert_ddp is not None:
            expert_ddp.reduce_gradients()
else:
    ddp = tdp.NaiveDDP(model, gradient_as_bucket_view=True, process_group=data_group)
    if world > 1:
        tdp.create_moe_dp_hooks(experts, dp_group, moe_dp_rank0)
    def finish():
        ddp.reduce_gradients()
        tdp.moe_dp_iter_step()

opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
gen = torch.Generator().manual_seed(100 + rank)
tok = torch.randint(0, cfg.vocab_size, (args.batch, args.seq + 1), generator=gen).to(dev)


def step():
    opt.zero_grad(set_to_none=False)
    loss = ddp(tok[:, :-1], tok[:, 1:])
    loss.backward()
    finish()
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
    print(json.dumps({"config": f"MoE {args.experts}-expert transformer EP={ep} x moe-DP={world // ep} "
                                f"(L={args.layers}, h={args.dim}, top-2, B={args.batch}/gpu, N={args.seq})",
                      "impl": args.impl, "n_gpus": world, "ms_per_step": t.item(),
                      "tokens_per_s": world * args.batch * args.seq / (t.item() / 1e3),
                      "dtype": "bf16", "loss": float(loss.item())}), flush=True)
dist.barrier()
dist.destroy_process_group()
