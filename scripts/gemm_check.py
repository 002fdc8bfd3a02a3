"""GPU check of the tcgen05 GEMM: numerics of every operand-major combination and epilogue against
an fp32 PyTorch reference, then timing vs cuBLAS (torch.matmul).  Run under gpurun."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchdistpackage_b200._C as C

torch.manual_seed(0)
dev = "cuda"
out = {"numerics": [], "timing": []}

def ref_mm(a, b, ta, tb):
    A = a.float().t() if ta else a.float()
    B = b.float().t() if tb else b.float()
    return A @ B

def check(M, N, K, ta, tb, block_n=0, **epi):
    a = torch.randn((K, M) if ta else (M, K), device=dev, dtype=torch.bfloat16)
    b = torch.randn((N, K) if tb else (K, N), device=dev, dtype=torch.bfloat16)
    c = torch.empty(M, N, device=dev, dtype=torch.float32 if epi.get("fp32") else torch.bfloat16)
    kw = {}
    ref = ref_mm(a, b, ta, tb)
    if epi.get("bias"):
        bias = torch.randn(N, device=dev, dtype=torch.bfloat16); kw["bias"] = bias
        ref = ref + bias.float()
    if epi.get("aux_out"):
        aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16); kw["aux_out"] = aux
        ref_aux = ref.clone()
    act = epi.get("act", 0)
    if act == 1: ref = torch.nn.functional.gelu(ref, approximate="tanh")
    if act == 2: ref = torch.nn.functional.gelu(ref)
    if act in (3, 4):
        z = torch.randn(M, N, device=dev, dtype=torch.bfloat16); kw["aux_in"] = z
        zz = z.float().requires_grad_(True)
        g = torch.nn.functional.gelu(zz, approximate="tanh" if act == 3 else "none")
        dg, = torch.autograd.grad(g.sum(), zz)
        ref = ref * dg
    if epi.get("residual"):
        res = torch.randn(M, N, device=dev, dtype=torch.bfloat16); kw["residual"] = res
        ref = ref + res.float()
    if epi.get("accumulate"):
        c.normal_(); ref = ref + c.float()
    C.gemm(a, b, c, ta, tb, act=act, accumulate=bool(epi.get("accumulate")), block_n=block_n, **kw)
    torch.cuda.synchronize()
    err = (c.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    rel = err / max(scale, 1e-6)
    rec = dict(M=M, N=N, K=K, ta=ta, tb=tb, block_n=block_n, epi=epi, max_abs_err=err, ref_max=scale, rel=rel, ok=bool(rel < 2e-2))
    if epi.get("aux_out"):
        rec["aux_rel"] = ((aux.float() - ref_aux).abs().max() / ref_aux.abs().max()).item()
        rec["ok"] = rec["ok"] and rec["aux_rel"] < 2e-2
    out["numerics"].append(rec)
    print(rec, flush=True)
    return rec["ok"]

ok = True
try:
    for ta in (False, True):
        for tb in (False, True):
            for bn in (128, 256):
                ok &= check(256, 512, 256, ta, tb, bn)
    ok &= check(1024, 2304, 768, False, False)
    ok &= check(1000, 760, 520, False, True)          # ragged edges (TMA OOB fill + masked stores)
    ok &= check(384, 3072, 768, False, False, bias=True, act=1, aux_out=True)
    ok &= check(384, 3072, 768, False, False, bias=True, act=2)
    ok &= check(384, 768, 3072, False, True, act=3)
    ok &= check(384, 768, 3072, False, True, act=4)
    ok &= check(512, 768, 3072, False, False, bias=True, residual=True)
    ok &= check(768, 3072, 2048, True, False, fp32=True, accumulate=True)
    ok &= check(768, 3072, 2048, True, False, accumulate=True)
    ok &= check(8192, 8192, 8192, False, False)
except Exception as e:
    import traceback; traceback.print_exc()
    out["error"] = repr(e); ok = False
out["all_ok"] = bool(ok)

def timeit(fn, iters=20, warm=5):
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]

if ok or os.environ.get("TDP_TIME_ANYWAY"):
    shapes = [(8192, 8192, 8192), (16384, 2304, 768), (16384, 768, 768), (16384, 3072, 768), (16384, 768, 3072),
              (8192, 1536, 4096), (8192, 4096, 512), (8192, 2048, 4096), (8192, 4096, 2048), (16384, 50304, 768)]
    for (M, N, K) in shapes:
        for (ta, tb) in ((False, False), (False, True), (True, False)):
            a = torch.randn((K, M) if ta else (M, K), device=dev, dtype=torch.bfloat16)
            b = torch.randn((N, K) if tb else (K, N), device=dev, dtype=torch.bfloat16)
            c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            A = a.t() if ta else a
            B = b.t() if tb else b
            t_ref = timeit(lambda: torch.matmul(A, B, out=c))
            rec = dict(M=M, N=N, K=K, ta=ta, tb=tb, cublas_ms=t_ref, cublas_tflops=2 * M * N * K / t_ref / 1e9)
            for bn in (128, 256):
                t = timeit(lambda: C.gemm(a, b, c, ta, tb, block_n=bn))
                rec[f"tdp{bn}_ms"] = t; rec[f"tdp{bn}_tflops"] = 2 * M * N * K / t / 1e9
            out["timing"].append(rec)
            print(rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_check.json", "w"), indent=1)
print("ALL_OK", ok)
