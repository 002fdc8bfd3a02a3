"""Grouped expert MLP (ops/grouped.py, one launch per product over all experts) vs the per-expert
loop on the same kernels: forward and every gradient, plus timing.  One B200:
    python scripts/grouped_check.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdistpackage_b200.ops import grouped, linear as L  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
res, ok = {"cases": []}, True


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


def loop_mlp(x, w1, b1, w2, b2):
    E = w1.shape[0]
    R = x.shape[0] // E
    return torch.cat([L.mlp(x[e * R:(e + 1) * R], w1[e], b1[e], w2[e], b2[e], layout="kn", act="gelu_tanh")
                      for e in range(E)])


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for (E, R, dim, hidden) in [(2, 128, 256, 512), (4, 256, 512, 1024), (2, 2560, 1024, 4096), (8, 640, 1024, 4096)]:
    mk = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).to(torch.bfloat16).requires_grad_(True)
    x, w1, b1 = mk(E * R, dim), mk(E, dim, hidden, sc=0.03), mk(E, hidden, sc=0.1)
    w2, b2 = mk(E, hidden, dim, sc=0.03), mk(E, dim, sc=0.1)
    gy = torch.randn(E * R, dim, device=dev).to(torch.bfloat16)
    assert grouped.grouped_supported(x, w1, w2)
    try:
        yg = grouped.grouped_mlp(x, w1, b1, w2, b2)
        gg = torch.autograd.grad(yg, (x, w1, b1, w2, b2), gy)
        yl = loop_mlp(x, w1, b1, w2, b2)
        gl = torch.autograd.grad(yl, (x, w1, b1, w2, b2), gy)
        torch.cuda.synchronize()
        errs = dict(y=rel(yg, yl), **{n: rel(a, b) for n, a, b in zip(("dx", "dw1", "db1", "dw2", "db2"), gg, gl)})
        good = max(errs.values()) < 3e-2
        t_g = timeit(lambda: torch.autograd.grad(grouped.grouped_mlp(x, w1, b1, w2, b2), (x, w1, w2), gy))
        t_l = timeit(lambda: torch.autograd.grad(loop_mlp(x, w1, b1, w2, b2), (x, w1, w2), gy))
    except Exception as ex:
        errs, good, t_g, t_l = {"error": repr(ex)}, False, None, None
    ok &= good
    rec = dict(E=E, R=R, dim=dim, hidden=hidden, ok=good, grouped_ms=t_g, loop_ms=t_l, **errs)
    res["cases"].append(rec)
    print(rec, flush=True)
res["all_ok"] = bool(ok)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/grouped_check.json", "w"), indent=1)
print("ALL_OK", ok)
