"""Tiny driver for Nsight Compute captures of the attention kernels (GPT-2 small shape):
    ncu --set full --clock-control none --import-source on -k regex:attn_ -s 3 -c 3 \
        -o gpurun_out/prof_attn python scripts/attn_prof.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TDP_ATTN"] = "native"
from torchdistpackage_b200.ops.attention import packed_attention  # noqa: E402

B, T, H = 16, 1024, 12
causal = os.environ.get("CAUSAL", "1") == "1"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
qkv = torch.randn(B, T, 3 * H * 64, device=dev).to(torch.bfloat16).requires_grad_(True)
dout = torch.randn(B, T, H * 64, device=dev).to(torch.bfloat16)
for _ in range(int(os.environ.get("ITERS", "3"))):
    o = packed_attention(qkv, H, causal)
    torch.autograd.grad(o, qkv, dout)
torch.cuda.synchronize()
print("done")
