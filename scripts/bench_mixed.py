"""BASELINE.json config #5: GPT-2 medium, DP x PP=2 x TP=2 (+sequence parallel), ZeRO over the data
group, 1F1B pipeline -- tokens/s device-timed (max over ranks).

  --impl ours       GPT2PipelineStage (fused TP/SP GEMM+collective kernels, flash attention),
                    forward_backward (NCCL p2p on a side stream), Bf16ZeroOptimizer (NVLS RS/AG +
                    fused Adam)
  --impl reference  the unmodified reference's ParallelBlock / forward_backward /
                    Bf16ZeroOptimizer / tpc, with plain-torch embeddings + LM head added by the
                    harness (the reference has no GPT model).  Attention is bidirectional in both
                    arms (the reference block has no mask).
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
ap.add_argument("--model", default="medium", choices=["medium", "small", "tiny"])
ap.add_argument("--batch", type=int, default=16, help="mini-batch per data-parallel replica")
ap.add_argument("--micro", type=int, default=4, help="micro-batches per step")
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()

if args.impl == "reference":
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import torchdistpackage as pkg
    from torchdistpackage.parallel import forward_backward
    from torchdistpackage.parallel.tensor_parallel.transformer import ParallelBlock as RefBlock
    from torchdistpackage.parallel.tensor_parallel import tp_utils as ref_tp
    try:
        pkg.setup_distributed("nccl")
    except UnboundLocalError:
        pass
else:
    import torchdistpackage_b200 as pkg
    from torchdistpackage_b200.parallel import forward_backward
    pkg.setup_distributed("nccl")
    pkg.tpc.verbose = False

rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", torch.cuda.current_device())
assert world % 4 == 0, "config #5 needs PP=2 x TP=2 (world multiple of 4)"
dp = world // 4
pkg.tpc.setup_process_groups([("data", dp), ("pipe", 2), ("tensor", 2)])
tpg, dpg = pkg.tpc.get_group("tensor"), pkg.tpc.get_group("data")
pp_rank = pkg.tpc.get_group_rank("pipe")
first, last = pp_rank == 0, pp_rank == 1

dims = {"medium": (24, 16, 1024, 1024, 50304), "small": (12, 12, 768, 1024, 50304),
        "tiny": (4, 4, 256, 256, 1024)}[args.model]
L, H, D, S, V = dims
torch.manual_seed(0)

if args.impl == "ours":
    from torchdistpackage_b200.models.gpt2 import GPT2Config
    from torchdistpackage_b200.models.gpt2_parallel import GPT2PipelineStage
    cfg = GPT2Config(vocab_size=V, n_layer=L, n_head=H, d_model=D, seq_len=S)
    stage = GPT2PipelineStage(cfg, tp_group=tpg, sequence_parallel=True)
    for blk in stage.blocks:                   # bidirectional, like the reference block
        blk.attn.causal = False
    stage = stage.to(dev).to(torch.bfloat16)
    fwd_fn = stage.forward_fn()
    opt = pkg.Bf16ZeroOptimizer(torch.optim.AdamW(stage.parameters(), lr=1e-4), dp_group=dpg,
                                overlap_comm=False)
    post_backward = stage.allreduce_replicated_grads
else:
    ref_tp.set_tp_group(tpg)

    class RefStage(nn.Module):
        def __init__(self):
            super().__init__()
            if first:
                self.wte, self.wpe = nn.Embedding(V, D), nn.Embedding(S, D)
            self.blocks = nn.ModuleList([RefBlock(D, mlp_ratio=4, num_heads=H, sequence_parallel=True)
                                         for _ in range(L // 2)])
            if last:
                self.ln_f = nn.LayerNorm(D)
                self.head = nn.Linear(D, V, bias=False)

    stage = RefStage()
    with torch.no_grad():
        g = torch.Generator().manual_seed(1)
        for p in stage.parameters():
            if p.dim() == 2:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    stage = stage.to(dev).to(torch.bfloat16)
    tp_rank = dist.get_rank(tpg)

    def fwd_fn(inp):
        items = list(inp) if isinstance(inp, (list, tuple)) else [inp]
        x = None if first else items.pop(0)
        if first:
            tok = items.pop(0)
            x = stage.wte(tok) + stage.wpe(torch.arange(tok.shape[1], device=dev))
        else:
            ref_tp.set_sequence_parallel_attr(x)   # the pipe delivers sequence-parallel shards
        for blk in stage.blocks:
            x = blk(x)            # first block splits into sequence-parallel shards
        if not last:
            return x
        tgt = items.pop(0)
        k = tgt.shape[0] // 2
        logits = stage.head(stage.ln_f(x))
        return F.cross_entropy(logits.view(-1, V), tgt[tp_rank * k:(tp_rank + 1) * k].reshape(-1)) / 2
    note = ""
    if args.micro > 1:
        # The reference's Bf16ZeroOptimizer frees p.grad inside its backward hook, so a second
        # micro-batch crashes in copy2master_or_free (zero_optim.py:222, 'NoneType'.data): ZeRO +
        # gradient accumulation is unsupported there.  Its arm therefore runs the reference's
        # NaiveDDP(num_grad_acc_iter) + a plain (fused) AdamW -- less communication than ZeRO.
        # (hooks sit on the parameters, so fwd_fn keeps calling the bare stage modules)
        ddp_stage = pkg.NaiveDDP(stage, sync=False, process_group=dpg, num_grad_acc_iter=args.micro,
                                 dp_rank0=pkg.tpc.get_ranks_in_group("data")[0]) if dp > 1 else None
        opt = torch.optim.AdamW(stage.parameters(), lr=1e-4, fused=True)
        post_backward = (lambda: ddp_stage.reduce_gradients()) if dp > 1 else (lambda: None)
        note = " [reference arm: NaiveDDP + AdamW, its ZeRO cannot accumulate gradients]"
    else:
        opt = pkg.Bf16ZeroOptimizer(torch.optim.AdamW(stage.parameters(), lr=1e-4), dp_group=dpg)
        post_backward = lambda: None

note = globals().get("note", "")
gen = torch.Generator().manual_seed(100 + pkg.tpc.get_group_rank("data"))
tok = torch.randint(0, V, (args.batch, S + 1), generator=gen).to(dev)
tokens, targets = tok[:, :-1].contiguous(), tok[:, 1:].contiguous()
inputs = ([tokens] if first else []) + ([targets] if last else [])


def step():
    kw = {"static_shapes": True} if args.impl == "ours" else {}   # (the reference re-handshakes)
    out = forward_backward(opt, fwd_fn, None, inputs or None, num_microbatches=args.micro,
                           dtype=torch.bfloat16, **kw)
    post_backward()
    opt.step()
    return out


for _ in range(args.warmup):
    out = step()
torch.cuda.synchronize(); dist.barrier()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(args.steps):
    out = step()
e.record(); torch.cuda.synchronize()
t = torch.tensor([s.elapsed_time(e) / args.steps], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
lossv = torch.tensor([float(out.detach().float().item()) * 2 if last else 0.0], device=dev)
dist.all_reduce(lossv, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": f"GPT-2 {args.model} DP={dp} x PP=2 x TP=2 (+SP), ZeRO over data, 1F1B, "
                                f"{args.micro} micro-batches, batch {args.batch}/replica, seq {S}"
                                + (note if args.impl == "reference" else ""),
                      "impl": args.impl, "n_gpus": world, "ms_per_step": t.item(),
                      "tokens_per_s": dp * args.batch * S / (t.item() / 1e3), "dtype": "bf16",
                      "loss": lossv.item()}), flush=True)
dist.barrier()
dist.destroy_process_group()
