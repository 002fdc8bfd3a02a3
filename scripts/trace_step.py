"""One eager training step under torch.profiler -> per-launch list (name, grid, duration) in launch
order, written to gpurun_out/trace_<impl>.json (kernels only)."""
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
for _ in range(4):
    step(tok[:, :-1], tok[:, 1:])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(tok[:, :-1], tok[:, 1:])
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
path = f"gpurun_out/_trace_{impl}_full.json"
prof.export_chrome_trace(path)
tr = json.load(open(path))
ks = [e for e in tr["traceEvents"] if e.get("cat") == "kernel"]
ks.sort(key=lambda e: e["ts"])
out = [dict(name=e["name"][:90], dur_us=e["dur"], grid=e["args"].get("grid"), block=e["args"].get("block"),
            ts=e["ts"] - ks[0]["ts"]) for e in ks]
json.dump(out, open(f"gpurun_out/trace_{impl}.json", "w"))
os.remove(path)
span = (ks[-1]["ts"] + ks[-1]["dur"] - ks[0]["ts"]) / 1e3
busy = sum(e["dur"] for e in ks) / 1e3
print(f"{impl}: {len(ks)} kernels, span {span:.2f} ms, sum of kernel time {busy:.2f} ms")
g = [e for e in out if "gemm_bf16" in e["name"] or "nvjet" in e["name"]]
for e in g[:60]:
    print(f'{e["dur_us"]:9.1f} us  grid {e["grid"]}  {e["name"][:60]}')
