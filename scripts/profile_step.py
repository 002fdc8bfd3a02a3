"""Per-kernel GPU time of one eager training step (torch.profiler / CUPTI), both bench arms.
usage: python scripts/profile_step.py {ours|reference} [out.json]"""
import json, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

impl = sys.argv[1] if len(sys.argv) > 1 else "ours"
args = types.SimpleNamespace(model="small", micro_batch=16, no_graph=True, impl=impl)
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
if impl == "reference":
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_bench
    step, _, cfg = ref_bench.build_reference(args, dev, 1)
else:
    step, _, cfg = bench.build_ours(args, dev, 1)
tok = torch.randint(0, cfg.vocab_size, (16, cfg.seq_len + 1), device=dev)
for _ in range(3):
    step(tok[:, :-1], tok[:, 1:])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step(tok[:, :-1], tok[:, 1:])
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    dt = getattr(e, "device_time_total", None)
    if dt is None: dt = getattr(e, "cuda_time_total", 0)
    if e.device_type is not None and "CUDA" in str(e.device_type) and dt > 0:
        rows.append((dt / 3 / 1e3, e.count // 3, e.key[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"== {impl}: total kernel time {tot:.2f} ms/step over {sum(r[1] for r in rows)} launches/step")
for ms, cnt, name in rows[:45]:
    print(f"{ms:8.3f} ms {cnt:5d}x  {name}")
out = sys.argv[2] if len(sys.argv) > 2 else f"gpurun_out/profile_step_{impl}.json"
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(dict(impl=impl, total_ms=tot, kernels=[dict(ms=r[0], count=r[1], name=r[2]) for r in rows]), open(out, "w"), indent=1)
