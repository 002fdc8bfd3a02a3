"""One-shot GPU check of the module profiler's CUDA-event path (times resolved once at report time,
memory deltas, backward hooks) and of the NaN hooks on device tensors."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
import torchdistpackage_b200 as tdp
from torchdistpackage_b200.tools import register_nan_hooks

dev = "cuda"
m = nn.Sequential(nn.Linear(512, 2048), nn.GELU(), nn.Sequential(nn.Linear(2048, 2048), nn.Linear(2048, 512))).to(dev).bfloat16()
x = torch.randn(8192, 512, device=dev, dtype=torch.bfloat16)
rep = tdp.get_model_profile(m, (x,), sort=False, backward=True)
rows = {n: (mb, f, b) for lvl in rep.values() for n, mb, f, b in lvl}
assert rows["root"][1] > 0 and rows["root"][2] > 0, rows
# fc1: allocator growth (its 33.5 MB output) minus the output-vs-input activation delta = 8.4 MB;
# GELU keeps its 33.5 MB input for backward
assert abs(rows["0"][0] - 8.4) < 1 and abs(rows["1"][0] - 33.6) < 1, rows
infos = tdp.register_profile_hooks(m)
for _ in range(3):
    m(x)
r2 = tdp.report_prof(infos, sort=True, min_mem=1)
infos.remove()
hs = register_nan_hooks(m)
try:
    m(torch.full_like(x, float("nan")))
    raise SystemExit("NaN hooks did not fire")
except FloatingPointError:
    pass
print("PROFILER_GPU_OK", {k: tuple(round(v, 3) for v in t) for k, t in rows.items()})
