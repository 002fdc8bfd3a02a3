"""Multi-GPU check (torchrun, N ranks on one box) of symmetric memory + custom collectives +
fused GEMM-collective kernels against NCCL / cuBLAS, with device-side timing (max over ranks)."""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdistpackage_b200.dist.launch import setup_distributed
from torchdistpackage_b200.ops.symm import SymmGroup
import torchdistpackage_b200._C as C

rank, world, _, _ = setup_distributed("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
torch.manual_seed(1234 + rank)
res = {"world": world}

def log(*a):
    if rank == 0:
        print(*a, flush=True)

def max_over_ranks(ms):
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return max_over_ranks(s.elapsed_time(e) / iters)

def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()

sg = SymmGroup(None, disable_multicast=os.environ.get("TDP_NO_MC") == "1")
log("symm enabled:", sg.enabled, sg.reason)
buf = sg.alloc(1 << 30)
res["multicast"] = buf.has_multicast
log("multicast:", buf.has_multicast)
ok = True

# ------------------------------------------------------------------ all-reduce
for dtype in (torch.bfloat16, torch.float32):
    for n in (1024, 1 << 20, (13 << 20) + 8):
        x = torch.randn(n, device=dev).to(dtype)
        algos = (3, 2) if buf.has_multicast else (2,)
        if n * x.element_size() <= 256 * 1024:           # one-shot latency path (NVLS / P2P loads)
            algos = algos + ((4, 1) if buf.has_multicast else (1,))
        for algo in algos:
            v = buf.view(0, (n,), dtype)
            v.copy_(x)
            buf.all_reduce_(0, n, dtype, 1.0 / world, algo=algo)
            ref = x.float().clone(); dist.all_reduce(ref); ref /= world
            torch.cuda.synchronize()
            r = rel(v, ref)
            good = r < (1e-2 if dtype == torch.bfloat16 else 1e-5)
            ok &= good
            log(f"all_reduce {dtype} n={n} algo={algo} rel={r:.2e} {'OK' if good else 'FAIL'}")

# ------------------------------------------------------------------ reduce-scatter / all-gather
n = 1 << 22
x = torch.randn(world * n, device=dev).to(torch.bfloat16)
buf.view(0, (world * n,), torch.bfloat16).copy_(x)
out32 = torch.zeros(n, device=dev)
buf.reduce_scatter(0, n, torch.bfloat16, out32, scale=1.0 / world)
ref = x.float().clone(); dist.all_reduce(ref); ref = ref[rank * n:(rank + 1) * n] / world
torch.cuda.synchronize()
r = rel(out32, ref); ok &= r < 1e-2; log(f"reduce_scatter->fp32 rel={r:.2e}")
outb = torch.zeros(n, device=dev, dtype=torch.bfloat16)
buf.reduce_scatter(0, n, torch.bfloat16, outb, scale=1.0 / world)
torch.cuda.synchronize()
r = rel(outb, ref); ok &= r < 1e-2; log(f"reduce_scatter->bf16 rel={r:.2e}")

mine = torch.randn(n, device=dev).to(torch.bfloat16)
buf.all_gather(0, n * 2, mine)
gathered = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
torch.cuda.synchronize()
good = torch.equal(buf.view(0, (world * n,), torch.bfloat16), torch.cat(gathered))
ok &= good; log("all_gather equal:", good)

# ------------------------------------------------------------------ bandwidth vs NCCL
bw = []
for mb in (1, 8, 25, 64, 256):
    n = mb * (1 << 20) // 2
    t = torch.randn(n, device=dev).to(torch.bfloat16)
    t_nccl = timeit(lambda: dist.all_reduce(t))
    rec = {"MiB": mb, "nccl_ms": t_nccl}
    for name, algo in (("nvls", 3), ("p2p", 2)):
        if algo == 3 and not buf.has_multicast:
            continue
        for ctas in (16, 32, 64):
            ms = timeit(lambda: buf.all_reduce_(0, n, torch.bfloat16, 1.0, algo=algo, max_ctas=ctas))
            rec[f"{name}{ctas}_ms"] = ms
    # bus bandwidth GB/s = S/t * 2(n-1)/n
    for k in list(rec):
        if k.endswith("_ms"):
            rec[k.replace("_ms", "_busbw")] = mb * (1 << 20) / (rec[k] * 1e-3) * 2 * (world - 1) / world / 1e9
    bw.append(rec); log(rec)
res["allreduce_bw"] = bw

# ------------------------------------------------------------------ small-message latency vs NCCL
lat = []
for kib in (2, 16, 64, 256):
    n = kib * 1024 // 2
    t = torch.randn(n, device=dev).to(torch.bfloat16)
    rec = {"KiB": kib, "nccl_us": timeit(lambda: dist.all_reduce(t)) * 1e3}
    for name, algo in (("one_shot_nvls", 4), ("one_shot_p2p", 1), ("two_shot_nvls", 3)):
        if algo in (3, 4) and not buf.has_multicast:
            continue
        rec[name + "_us"] = timeit(lambda: buf.all_reduce_(0, n, torch.bfloat16, 1.0, algo=algo)) * 1e3
    lat.append(rec); log(rec)
res["allreduce_latency"] = lat

# ------------------------------------------------------------------ fused GEMM -> reduce-scatter
T, Kl, N = 2048 * world if world <= 4 else 8192, 2048, 4096
rows = T // world
a = (torch.randn(T, Kl, device=dev) * 0.5).to(torch.bfloat16)
w = (torch.randn(Kl, N, device=dev) * 0.05).to(torch.bfloat16)
stage_off = 0
cnt_word = buf.alloc_words(8)
units_per_call = (rows // 32) * (N // 8)   # counter units each src delivers per call
def run_gemm_rs(out):
    buf.barrier(1)
    buf.handle.gemm_rs(a, w, False, stage_off, cnt_word, 0)
    tgt = buf.next_epoch(cnt_word, units_per_call)
    buf.handle.rs_reduce(stage_off, rows, N, cnt_word, tgt, None, None, out, False, 0, True, 0)
out = torch.empty(rows, N, device=dev, dtype=torch.bfloat16)
try:
    run_gemm_rs(out)
    full = (a.float() @ w.float())
    dist.all_reduce(full)
    torch.cuda.synchronize()
    r = rel(out, full[rank * rows:(rank + 1) * rows]); good = r < 2e-2
    ok &= good; log(f"gemm_rs rel={r:.2e} {'OK' if good else 'FAIL'}")
    def ref_rs():
        y = a @ w
        o = torch.empty(rows, N, device=dev, dtype=torch.bfloat16)
        dist.reduce_scatter_tensor(o, y)
    t_ref = timeit(ref_rs); t_fused = timeit(lambda: run_gemm_rs(out))
    t_gemm = timeit(lambda: torch.matmul(a, w))
    res["gemm_rs"] = dict(T=T, K=Kl, N=N, cublas_nccl_ms=t_ref, fused_ms=t_fused, cublas_gemm_only_ms=t_gemm)
    log(res["gemm_rs"])
except Exception as ex:
    import traceback; traceback.print_exc(); ok = False; res["gemm_rs_error"] = repr(ex)

# ------------------------------------------------------------------ fused all-gather -> GEMM
Kf, Nf = 4096, 2048
xs = (torch.randn(rows, Kf, device=dev) * 0.5).to(torch.bfloat16)
w2 = (torch.randn(Kf, Nf, device=dev) * 0.05).to(torch.bfloat16)
ag_off = 512 << 20
flag_word = buf.alloc_words(8)
cfull = torch.empty(T, Nf, device=dev, dtype=torch.bfloat16)
def run_ag_gemm():
    buf.barrier(2)
    ep = buf.next_epoch(flag_word)
    buf.handle.all_gather_signal(ag_off, rows * Kf * 2, xs, flag_word, ep, True, 16)
    buf.handle.gemm_ag(ag_off, rows, Kf, w2, False, cfull, None, None, 0, flag_word, ep, 0, xs, None)
try:
    run_ag_gemm()
    gl = [torch.empty_like(xs) for _ in range(world)]
    dist.all_gather(gl, xs)
    refc = torch.cat(gl).float() @ w2.float()
    torch.cuda.synchronize()
    r = rel(cfull, refc); good = r < 2e-2
    ok &= good; log(f"ag_gemm rel={r:.2e} {'OK' if good else 'FAIL'}")
    xfull = torch.empty(T, Kf, device=dev, dtype=torch.bfloat16)
    def ref_ag():
        dist.all_gather_into_tensor(xfull, xs)
        torch.matmul(xfull, w2, out=cfull)
    t_ref = timeit(ref_ag); t_fused = timeit(run_ag_gemm)
    t_gemm = timeit(lambda: torch.matmul(xfull, w2, out=cfull))
    res["ag_gemm"] = dict(T=T, K=Kf, N=Nf, nccl_cublas_ms=t_ref, fused_ms=t_fused, cublas_gemm_only_ms=t_gemm)
    log(res["ag_gemm"])
except Exception as ex:
    import traceback; traceback.print_exc(); ok = False; res["ag_gemm_error"] = repr(ex)

# ------------------------------------------------------------------ all-gather *inside* the GEMM
def run_ag_gemm_inkernel():
    buf.barrier(2)
    ep = buf.next_epoch(flag_word)
    buf.handle.gemm_ag(ag_off, rows, Kf, w2, False, cfull, None, None, 0, flag_word, ep, 0, xs, None,
                       True, True)
try:
    buf.view(ag_off, (T, Kf), torch.bfloat16).zero_(); cfull.zero_()
    torch.cuda.synchronize(); dist.barrier()
    run_ag_gemm_inkernel()
    torch.cuda.synchronize(); dist.barrier()
    r = rel(cfull, refc)
    same = torch.equal(buf.view(ag_off, (T, Kf), torch.bfloat16), torch.cat(gl))
    good = r < 2e-2 and same
    ok &= good; log(f"ag_gemm in-kernel push rel={r:.2e} gathered_equal={same} {'OK' if good else 'FAIL'}")
    res["ag_gemm_inkernel"] = dict(T=T, K=Kf, N=Nf, fused_ms=timeit(run_ag_gemm_inkernel),
                                   nccl_cublas_ms=timeit(ref_ag))
    log(res["ag_gemm_inkernel"])
except Exception as ex:
    import traceback; traceback.print_exc(); ok = False; res["ag_gemm_inkernel_error"] = repr(ex)

# ------------------------------------------------------------------ a2a rows
hid, nrow = 1024, 4096
src = torch.randn(nrow, hid, device=dev).to(torch.bfloat16)
dst_rank = torch.randint(0, world, (nrow,), device=dev, dtype=torch.int32)
# destination row = (src_rank * nrow + i): unique slot per (src,row) on every destination
dst_row = (rank * nrow + torch.arange(nrow, device=dev)).to(torch.int32)
a2a_off = 768 << 20
buf.barrier(3)
buf.handle.a2a_scatter_rows(a2a_off, src, dst_rank, dst_row)
buf.barrier(3)
torch.cuda.synchronize()
# verify: gather back what we scattered
back = torch.empty_like(src)
buf.handle.a2a_gather_rows(a2a_off, back, dst_rank, dst_row, None, False)
torch.cuda.synchronize()
good = torch.equal(back, src); ok &= good; log("a2a scatter/gather roundtrip equal:", good)
buf.barrier(3)

res["all_ok"] = bool(ok)
allok = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(allok, op=dist.ReduceOp.MIN)
res["all_ranks_ok"] = bool(allok.item())
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/symm_check_w{world}.json", "w"), indent=1)
    print("ALL_OK", res["all_ranks_ok"], flush=True)
dist.barrier()
dist.destroy_process_group()
