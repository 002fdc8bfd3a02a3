"""torch.fx per-node timing and a time-balanced 2-way pipeline split (reference study:
explore/fx/torch_fx_profile.py, fx_graph_split.py).  Limitation kept from the reference: a cut may
only be placed where exactly one value crosses it (no residual edges across the cut)."""
import time
import torch, torch.nn as nn, torch.fx as fx


class ProfilingInterpreter(fx.Interpreter):
    def __init__(self, gm):
        super().__init__(gm)
        self.times = {}

    def run_node(self, n):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = super().run_node(n)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.times[n.name] = self.times.get(n.name, 0.0) + (time.perf_counter() - t0) * 1e3
        return out


def live_values_after(gm: fx.GraphModule):
    """For every node index i: names produced at or before i and used after i."""
    nodes = list(gm.graph.nodes)
    last_use = {}
    for i, n in enumerate(nodes):
        for a in n.all_input_nodes:
            last_use[a.name] = i
    live = []
    for i, n in enumerate(nodes):
        live.append({m.name for m in nodes[:i + 1] if last_use.get(m.name, -1) > i
                     and m.op != "get_attr"})
    return nodes, live


def balanced_cut(gm: fx.GraphModule, times: dict):
    nodes, live = live_values_after(gm)
    total = sum(times.get(n.name, 0.0) for n in nodes)
    best, acc = None, 0.0
    for i, n in enumerate(nodes[:-1]):
        acc += times.get(n.name, 0.0)
        if len(live[i]) == 1 and n.op not in ("placeholder",):
            score = abs(acc - total / 2)
            if best is None or score < best[0]:
                best = (score, i, acc)
    return best, total


if __name__ == "__main__":
    model = nn.Sequential(nn.Linear(256, 1024), nn.GELU(), nn.Linear(1024, 1024), nn.GELU(),
                          nn.Linear(1024, 256), nn.LayerNorm(256))
    gm = fx.symbolic_trace(model)
    x = torch.randn(512, 256)
    interp = ProfilingInterpreter(gm)
    for _ in range(3):
        interp.run(x)
    for k, v in interp.times.items():
        print(f"{k:24s} {v / 3:8.3f} ms")
    best, total = balanced_cut(gm, interp.times)
    print(f"total {total / 3:.3f} ms; cut after node #{best[1]} with {best[2] / 3:.3f} ms on stage 0")
