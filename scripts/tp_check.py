"""N-GPU equivalence of the tensor+sequence-parallel transformer block (fused GEMM+collective
kernels) against the serial block; then step timing fused vs NCCL path."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchdistpackage_b200 as tdp
from torchdistpackage_b200.parallel.tensor_parallel import tp_fused
from torchdistpackage_b200.parallel.tensor_parallel.transformer import (
    Block, ParallelBlock, Transformer, allreduce_sequence_parallel_grads)

rank, world, _, _ = tdp.setup_distributed("nccl")
tdp.tpc.verbose = False
tdp.tpc.setup_process_groups([("tensor", world)])
dev = torch.device("cuda", torch.cuda.current_device())
res = {"world": world}
def log(*a):
    if rank == 0: print(*a, flush=True)
def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()

dim, heads, B, N = 1024, 16, 2 * world, 512
torch.manual_seed(7)
serial = Block(dim, num_heads=heads).to(dev)
with torch.no_grad():
    for n, p in serial.named_parameters():
        if p.dim() == 2: p.copy_((torch.rand_like(p) - 0.5) * 0.08)
        elif "bias" in n: p.copy_((torch.rand_like(p) - 0.5) * 0.1)
for p in serial.parameters():
    dist.broadcast(p.data, 0)
serial = serial.to(torch.bfloat16)
x = torch.randn(B, N, dim, device=dev).to(torch.bfloat16)
dist.broadcast(x, 0)
gy = torch.randn(B, N, dim, device=dev).to(torch.bfloat16)
dist.broadcast(gy, 0)

xs = x.clone().requires_grad_(True)
ys = serial(xs); ys.backward(gy)

ok = True
for fused in (True, False):
    tp_fused.set_enabled(fused)
    par = ParallelBlock(dim, num_heads=heads, sequence_parallel=True).to(dev).to(torch.bfloat16)
    par.init_from_full(serial)
    xp = x.clone().requires_grad_(True)
    yp = par(xp)                      # [B/world, N, dim] shard of dim 0
    k = B // world
    yp.backward(gy[rank * k:(rank + 1) * k])
    allreduce_sequence_parallel_grads(par)
    torch.cuda.synchronize()
    r_fwd = rel(yp, ys[rank * k:(rank + 1) * k])
    r_dx = rel(xp.grad[rank * k:(rank + 1) * k], xs.grad[rank * k:(rank + 1) * k])
    h = 4 * dim // world
    r_w1 = rel(par.mlp.fc1.linear.weight.grad, serial.mlp.fc1.weight.grad[:, rank * h:(rank + 1) * h])
    r_w2 = rel(par.mlp.fc2.linear.weight.grad, serial.mlp.fc2.weight.grad[rank * h:(rank + 1) * h])
    d = dim // world
    r_proj = rel(par.attn.proj.linear.weight.grad, serial.attn.proj.weight.grad[rank * d:(rank + 1) * d])
    r_ln = rel(par.ln_2.weight.grad, serial.ln_2.weight.grad)
    r_b2 = rel(par.mlp.fc2.linear.bias.grad, serial.mlp.fc2.bias.grad)
    rec = dict(fused=fused, fwd=r_fwd, dx=r_dx, dw_fc1=r_w1, dw_fc2=r_w2, dw_proj=r_proj, dln2=r_ln, db2=r_b2)
    good = all(v < 4e-2 for k_, v in rec.items() if k_ != "fused")
    ok &= good
    res[f"block_fused_{fused}"] = rec
    log(rec, "OK" if good else "FAIL")

# ---- plain tensor parallel (no SP): row-parallel proj runs the fused GEMM -> all-reduce
for fused in (True, False):
    tp_fused.set_enabled(fused)
    par = ParallelBlock(dim, num_heads=heads, sequence_parallel=False).to(dev).to(torch.bfloat16)
    par.init_from_full(serial)
    xp = x.clone().requires_grad_(True)
    yp = par(xp)
    yp.backward(gy)
    torch.cuda.synchronize()
    d = dim // world
    rec = dict(fused=fused, fwd=rel(yp, ys), dx=rel(xp.grad, xs.grad),
               dw_proj=rel(par.attn.proj.linear.weight.grad, serial.attn.proj.weight.grad[rank * d:(rank + 1) * d]),
               db_proj=rel(par.attn.proj.linear.bias.grad, serial.attn.proj.bias.grad))
    good = all(v < 4e-2 for k_, v in rec.items() if k_ != "fused")
    ok &= good
    res[f"block_nosp_fused_{fused}"] = rec
    log("no-SP", rec, "OK" if good else "FAIL")

# ---- timing: config #3-like transformer (4 blocks), fused vs NCCL path
def step_time(fused, dim=4096, heads=32, depth=4, B=4, N=2048, iters=6, warm=3):
    tp_fused.set_enabled(fused)
    torch.manual_seed(3)
    model = Transformer(dim, num_heads=heads, depth=depth, tensor_parallel=True, sequence_parallel=True)
    for m in model.modules():
        if hasattr(m, "reset_parameters_scaled"): m.reset_parameters_scaled()
    model = model.to(dev).to(torch.bfloat16)
    xin = torch.randn(B, N, dim, device=dev).to(torch.bfloat16)
    def one():
        out = model(xin)
        out.float().mean().backward()
        allreduce_sequence_parallel_grads(model)
    for _ in range(warm): one()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(iters): one()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del model
    torch.cuda.empty_cache()
    return t.item()
try:
    if True:
        Bt = max(4, world)          # sequence parallelism shards dim 0: needs B % tp == 0
        t_f = step_time(True, B=Bt); t_n = step_time(False, B=Bt)
        res["tp_transformer_ms"] = dict(fused=t_f, nccl=t_n, tokens=Bt * 2048)
        log("transformer fwd+bwd ms: fused", t_f, "nccl", t_n)
except Exception as ex:
    import traceback; traceback.print_exc(); res["timing_error"] = repr(ex)

res["all_ok"] = bool(ok)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/tp_check_w{world}.json", "w"), indent=1)
    print("ALL_OK", ok, flush=True)
dist.barrier(); dist.destroy_process_group()
