"""Run ONE GEMM configuration a few times (for ncu):
   python scripts/gemm_one.py M N K ta tb cta_group block_n [act]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdistpackage_b200.ops._loader import native  # noqa: E402

C = native(required=True)
M, N, K, ta, tb, cg, bn = [int(v) for v in sys.argv[1:8]]
act = int(sys.argv[8]) if len(sys.argv) > 8 else 0
dev = torch.device("cuda", 0)
a = torch.randn((K, M) if ta else (M, K), device=dev, dtype=torch.bfloat16)
b = torch.randn((N, K) if tb else (K, N), device=dev, dtype=torch.bfloat16)
c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
kw = {}
if act == 3:
    kw["aux_in"] = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
if act == 1:
    kw["aux_out"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    kw["bias"] = torch.zeros(N, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for _ in range(4):
    flush.zero_()
    C.gemm(a, b, c, bool(ta), bool(tb), act=act, cta_group=cg, block_n=bn, **kw)
torch.cuda.synchronize()
print("done", M, N, K, ta, tb, cg, bn, act)
