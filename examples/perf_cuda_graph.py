"""Eager vs CUDA-graph replay of a whole training step (reference study: explore/perf/test_timm.py,
which tried AMP / AOT fusion / CUDA graphs on timm models).  The step of a small GPT-2 is
launch-bound at small batch; capturing it removes the CPU launch cost."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchdistpackage_b200 as tdp
from torchdistpackage_b200.models.gpt2 import build_gpt2

assert torch.cuda.is_available(), "needs a GPU"
dev = torch.device("cuda")
model = build_gpt2("tiny", device=dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, capturable=True)
tok = torch.randint(0, model.cfg.vocab_size, (2, model.cfg.seq_len + 1), device=dev)

def step():
    opt.zero_grad(set_to_none=False)
    loss = model(tok[:, :-1], tok[:, 1:]); loss.backward(); opt.step()
    return loss

def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

eager = timed(step)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    static_loss = step()
graphed = timed(g.replay)
print(f"eager {eager:.3f} ms/step, CUDA graph {graphed:.3f} ms/step, loss {static_loss.item():.3f}")
