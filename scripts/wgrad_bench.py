"""Weight-gradient shaped GEMMs (dW = X^T dY, few output tiles, K = tokens): tile / split-K sweep
vs cuBLAS.  `--one` runs a single configuration a few times (for ncu)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchdistpackage_b200._C as C

dev = "cuda"
def timeit(fn, iters=20, warm=5):
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]

if "--one" in sys.argv:
    K, M, N = 16384, 768, 2304
    x = torch.randn(K, M, device=dev, dtype=torch.bfloat16); dy = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
    acc = torch.zeros(M, N, device=dev)
    for _ in range(6):
        C.gemm(x, dy, acc, True, False, split_k=4, block_n=128)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        C.gemm(x, dy, out, True, False, split_k=1, block_n=256)
    torch.cuda.synchronize(); sys.exit(0)

res = []
for (K, M, N) in [(16384, 768, 2304), (16384, 768, 768), (16384, 768, 3072), (16384, 3072, 768), (8192, 4096, 1536), (8192, 512, 4096)]:
    x = torch.randn(K, M, device=dev, dtype=torch.bfloat16); dy = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); acc = torch.zeros(M, N, device=dev)
    flops = 2.0 * M * N * K
    rec = dict(K=K, M=M, N=N, cublas_tflops=flops / timeit(lambda: torch.matmul(x.t(), dy, out=out)) / 1e9)
    # same product with K-major operands (pre-transposed copies) to isolate the MN-major path
    xt, dyt = x.t().contiguous(), dy.t().contiguous()
    rec["tdp_kmajor_bn256"] = flops / timeit(lambda: C.gemm(xt, dyt, out, False, True, block_n=256)) / 1e9
    for bn in (128, 256):
        rec[f"tdp_bn{bn}_s1"] = flops / timeit(lambda: C.gemm(x, dy, out, True, False, block_n=bn)) / 1e9
        for s in (2,):
            rec[f"tdp_bn{bn}_s{s}"] = flops / timeit(lambda: C.gemm(x, dy, acc, True, False, block_n=bn, split_k=s)) / 1e9
    res.append(rec); print({k: (round(v) if isinstance(v, float) else v) for k, v in rec.items()}, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/wgrad_bench.json", "w"), indent=1)
