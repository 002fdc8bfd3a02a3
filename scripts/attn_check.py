"""tcgen05 flash-attention forward (csrc/attn/attn_fwd_sm100.cu) vs an fp32 reference and vs the
library kernel: numerics (output + LSE, causal and full) and timing.  One B200:
    python scripts/attn_check.py"""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdistpackage_b200.ops.attention import native_attention_forward  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
res = {"numerics": [], "timing": []}
ok = True


def ref_attn(qkv, H, causal):
    B, T, D3 = qkv.shape
    dh = D3 // 3 // H
    q, k, v = qkv.float().view(B, T, 3, H, dh).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * dh ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(T, T, device=dev, dtype=torch.bool).triu(1), float("-inf"))
    lse = torch.logsumexp(s, -1)
    o = torch.softmax(s, -1) @ v
    return o.transpose(1, 2).reshape(B, T, H * dh), lse


for (B, T, H, causal) in [(1, 128, 1, False), (1, 128, 1, True), (2, 256, 2, True), (2, 384, 3, True),
                          (2, 512, 4, False), (4, 1024, 12, True)]:
    qkv = (torch.randn(B, T, 3 * H * 64, device=dev) * 0.7).to(torch.bfloat16)
    try:
        out, lse = native_attention_forward(qkv, H, causal, return_lse=True)
        torch.cuda.synchronize()
        ro, rl = ref_attn(qkv, H, causal)
        e_o = ((out.float() - ro).abs().max() / ro.abs().max()).item()
        e_l = (lse - rl).abs().max().item()
        good = e_o < 2e-2 and e_l < 2e-2
    except Exception as ex:      # keep going: the first failing shape is the interesting one
        e_o = e_l = float("nan"); good = False
        print("ERROR", repr(ex), flush=True)
    ok &= good
    rec = dict(B=B, T=T, H=H, causal=causal, out_rel=e_o, lse_abs=e_l, ok=good)
    res["numerics"].append(rec)
    print(rec, flush=True)
    if not good:
        break


# ---- backward: dqkv of the native kernels vs autograd through the fp32 reference
os.environ["TDP_ATTN"] = "native"
from torchdistpackage_b200.ops.attention import packed_attention  # noqa: E402
if ok:
    for (B, T, H, causal) in [(1, 128, 1, False), (1, 256, 1, True), (2, 384, 2, True), (2, 512, 4, False),
                              (2, 1024, 12, True)]:
        qkv = (torch.randn(B, T, 3 * H * 64, device=dev) * 0.7).to(torch.bfloat16).requires_grad_(True)
        dout = (torch.randn(B, T, H * 64, device=dev) * 0.5).to(torch.bfloat16)
        try:
            out = packed_attention(qkv, H, causal)
            out.backward(dout)
            torch.cuda.synchronize()
            qr = qkv.detach().float().requires_grad_(True)
            ro, _ = ref_attn(qr, H, causal)
            ro.backward(dout.float())
            g, gr = qkv.grad.float().view(B, T, 3, H, 64), qr.grad.view(B, T, 3, H, 64)
            errs = [((g[:, :, i] - gr[:, :, i]).abs().max() / gr[:, :, i].abs().max()).item() for i in range(3)]
            good = max(errs) < 3e-2
        except Exception as ex:
            errs = [float("nan")] * 3; good = False
            print("ERROR", repr(ex), flush=True)
        ok &= good
        rec = dict(bwd=True, B=B, T=T, H=H, causal=causal, dq_rel=errs[0], dk_rel=errs[1], dv_rel=errs[2], ok=good)
        res["numerics"].append(rec)
        print(rec, flush=True)
        if not good:
            break


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


if ok:
    B, T, H = 16, 1024, 12
    qkv = torch.randn(B, T, 3 * H * 64, device=dev).to(torch.bfloat16)
    q, k, v = qkv.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    for causal in (True, False):
        fl = 4 * B * H * T * T * 64 / (2 if causal else 1)
        t_lib = timeit(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=causal))
        t_own = timeit(lambda: native_attention_forward(qkv, H, causal))
        rec = dict(B=B, T=T, H=H, causal=causal, library_ms=t_lib, native_ms=t_own,
                   library_tflops=fl / t_lib / 1e9, native_tflops=fl / t_own / 1e9)
        res["timing"].append(rec)
        print(rec, flush=True)
        # forward + backward
        qg = qkv.clone().requires_grad_(True)
        q4, k4, v4 = (t.detach().requires_grad_(True) for t in (q, k, v))
        dout = torch.randn(B, T, H * 64, device=dev).to(torch.bfloat16)
        do_lib = dout.view(B, T, H, 64).transpose(1, 2)

        def lib_step():
            o = F.scaled_dot_product_attention(q4, k4, v4, is_causal=causal)
            torch.autograd.grad(o, (q4, k4, v4), do_lib)

        def own_step():
            o = packed_attention(qg, H, causal)
            torch.autograd.grad(o, qg, dout)
        rec2 = dict(B=B, T=T, H=H, causal=causal, fwd_bwd=True, library_ms=timeit(lib_step),
                    native_ms=timeit(own_step))
        res["timing"].append(rec2)
        print(rec2, flush=True)
res["all_ok"] = bool(ok)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/attn_check.json", "w"), indent=1)
print("ALL_OK", ok)
