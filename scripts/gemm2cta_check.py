"""Numerics + timing of the 2-CTA (cta_group::2, 256x256 per CTA pair) GEMM against an fp32
reference, the 1-CTA kernel and cuBLAS.  Run on one B200:  python scripts/gemm2cta_check.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdistpackage_b200.ops._loader import native  # noqa: E402

C = native(required=True)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
out = {"numerics": [], "timing": []}


def check(M, N, K, ta, tb, block_n=0, **epi):
    a = torch.randn((K, M) if ta else (M, K), device=dev, dtype=torch.bfloat16)
    b = torch.randn((N, K) if tb else (K, N), device=dev, dtype=torch.bfloat16)
    c = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    ref = (a.float().t() if ta else a.float()) @ (b.float().t() if tb else b.float())
    kw = {}
    if epi.get("bias"):
        kw["bias"] = torch.randn(N, device=dev, dtype=torch.bfloat16)
        ref = ref + kw["bias"].float()
    if epi.get("aux_out"):
        kw["aux_out"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ref_aux = ref.clone()
    act = epi.get("act", 0)
    if act == 1:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    if act == 3:
        z = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        kw["aux_in"] = z
        zz = z.float().requires_grad_(True)
        dg, = torch.autograd.grad(torch.nn.functional.gelu(zz, approximate="tanh").sum(), zz)
        ref = ref * dg
    if epi.get("residual"):
        kw["residual"] = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        ref = ref + kw["residual"].float()
    C.gemm(a, b, c, ta, tb, act=act, cta_group=2, block_n=block_n, **kw)
    torch.cuda.synchronize()
    rel = ((c.float() - ref).abs().max() / ref.abs().max()).item()
    rec = dict(M=M, N=N, K=K, ta=ta, tb=tb, block_n=block_n, epi=epi, rel=rel, ok=bool(rel < 2e-2))
    if epi.get("aux_out"):
        rec["aux_rel"] = ((kw["aux_out"].float() - ref_aux).abs().max() / ref_aux.abs().max()).item()
        rec["ok"] = rec["ok"] and rec["aux_rel"] < 2e-2
    out["numerics"].append(rec)
    print(rec, flush=True)
    return rec["ok"]


ok = True
for ta in (False, True):
    for tb in (False, True):
        ok &= check(512, 512, 256, ta, tb)
        ok &= check(512, 512, 256, ta, tb, block_n=128)
ok &= check(1024, 2304, 768, False, False)
ok &= check(768, 3072, 2048, True, False, block_n=128)
ok &= check(1000, 760, 520, False, True, block_n=128)
ok &= check(384, 3072, 768, False, False, block_n=128, bias=True, act=1, aux_out=True)
ok &= check(1000, 760, 520, False, True)
ok &= check(1000, 760, 520, True, False)
ok &= check(384, 3072, 768, False, False, bias=True, act=1, aux_out=True)
ok &= check(384, 768, 3072, False, True, act=3)
ok &= check(512, 768, 3072, False, False, bias=True, residual=True)
ok &= check(8192, 8192, 8192, False, True)
out["all_ok"] = bool(ok)


def timeit(fn, iters=20, warm=5):
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


if ok:
    shapes = [(8192, 8192, 8192), (16384, 2304, 768), (16384, 768, 768), (16384, 3072, 768),
              (16384, 768, 3072), (16384, 768, 2304), (8192, 4096, 2048), (16384, 50304, 768),
              (16384, 768, 50304)]
    for (M, N, K) in shapes:
        for (ta, tb) in ((False, False), (False, True)):
            a = torch.randn((K, M) if ta else (M, K), device=dev, dtype=torch.bfloat16)
            b = torch.randn((N, K) if tb else (K, N), device=dev, dtype=torch.bfloat16)
            c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            A = a.t() if ta else a
            B = b.t() if tb else b
            fl = 2 * M * N * K / 1e9
            t0 = timeit(lambda: torch.matmul(A, B, out=c))
            t1 = timeit(lambda: C.gemm(a, b, c, ta, tb, cta_group=1))
            t2 = timeit(lambda: C.gemm(a, b, c, ta, tb, cta_group=2))
            rec = dict(M=M, N=N, K=K, ta=ta, tb=tb, cublas_ms=t0, tdp_1cta_ms=t1, tdp_2cta_ms=t2,
                       cublas_tflops=fl / t0, tdp_1cta_tflops=fl / t1, tdp_2cta_tflops=fl / t2)
            out["timing"].append(rec)
            print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in rec.items()}, flush=True)
    # weight-gradient shapes (x^T @ dy, K = tokens): stream-K 1-CTA (fp32 atomics + cast) vs
    # 256x128 / 256x256 CTA-pair tiles writing bf16 directly
    from torchdistpackage_b200.ops.linear import gemm as tdp_gemm
    for (M, N, K) in [(768, 3072, 16384), (3072, 768, 16384), (768, 2304, 16384), (768, 768, 16384),
                      (50304, 768, 16384)]:
        a = torch.randn(K, M, device=dev, dtype=torch.bfloat16)
        b = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
        c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fl = 2 * M * N * K / 1e9
        rec = dict(wgrad=True, M=M, N=N, K=K)
        rec["cublas_ms"] = timeit(lambda: torch.matmul(a.t(), b, out=c))
        rec["tdp_streamk_ms"] = timeit(lambda: tdp_gemm(a, b, trans_a=True))
        rec["tdp_1cta_ms"] = timeit(lambda: C.gemm(a, b, c, True, False, cta_group=1))
        rec["tdp_2cta_n128_ms"] = timeit(lambda: C.gemm(a, b, c, True, False, cta_group=2, block_n=128))
        rec["tdp_2cta_n256_ms"] = timeit(lambda: C.gemm(a, b, c, True, False, cta_group=2, block_n=256))
        out["timing"].append(rec)
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in rec.items()}, flush=True)
    # fused epilogues (the shapes of the GPT-2 MLP at 16k tokens)
    M, H = 16384, 768
    x = torch.randn(M, H, device=dev, dtype=torch.bfloat16)
    w1 = torch.randn(4 * H, H, device=dev, dtype=torch.bfloat16) * 0.02
    b1 = torch.zeros(4 * H, device=dev, dtype=torch.bfloat16)
    h = torch.empty(M, 4 * H, device=dev, dtype=torch.bfloat16)
    z = torch.empty(M, 4 * H, device=dev, dtype=torch.bfloat16)
    dy = torch.randn(M, H, device=dev, dtype=torch.bfloat16)
    w2 = torch.randn(H, 4 * H, device=dev, dtype=torch.bfloat16) * 0.02
    for g in (1, 2):
        rec = dict(cta_group=g)
        rec["fc1_gelu_aux_ms"] = timeit(lambda: C.gemm(x, w1, h, False, True, bias=b1, act=1, aux_out=z, cta_group=g))
        rec["dgelu_ms"] = timeit(lambda: C.gemm(dy, w2, h, False, False, act=3, aux_in=z, cta_group=g))
        rec["proj_res_ms"] = timeit(lambda: C.gemm(h, w2, dy.clone(), False, True, residual=x, cta_group=g))
        out["timing"].append(rec)
        print(rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm2cta_check.json", "w"), indent=1)
print("ALL_OK", ok)
