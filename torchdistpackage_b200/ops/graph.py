"""CUDA-graph capture of a whole training step.

The framework's answer to launch-bound steps is streams + graphs, not a tracing compiler: the
eager step (forward, backward with the DDP bucket reductions forked onto the comm stream,
fused optimizer) is warmed up on a side stream, captured once, and replayed.  Everything on the
hot path is capture-safe by construction:

* the NVLS collectives synchronise with stateless in-kernel barriers (no host-side epochs),
* BucketAdamW keeps ``step`` / ``lr`` in device memory and advances them inside the graph,
* TMA tensor maps are kernel parameters (baked at capture; buffers are static in the graph pool).

(The fused sequence-parallel GEMM+collective ops carry host-incremented epoch flags and are not
graph-captured yet.)
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    """``g = GraphedStep(step_fn, (tokens, targets)); loss = g(tokens, targets)``.

    ``step_fn`` must be a pure function of its tensor arguments plus state that lives in device
    memory (parameters, optimizer state).  The returned tensors are static: read them (``.item()``)
    before the next replay."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3,
                 preserve: Sequence[torch.Tensor] = ()):
        """``preserve``: tensors the step mutates (parameters, optimizer state -- e.g.
        ``list(model.parameters()) + opt.state_tensors()``) that must come out of the warm-up
        unchanged: they are snapshotted before the warm-up iterations (which are *real* steps on
        the example inputs) and restored before capture, so building the graph after a
        checkpoint resume does not advance the run."""
        assert torch.cuda.is_available()
        self.static_inputs = [t.clone() for t in example_inputs]
        saved = [(t, t.detach().clone()) for t in preserve]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_inputs)
        cur.wait_stream(side)
        with torch.no_grad():
            for t, snap in saved:
                t.copy_(snap)
        del saved
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_outputs = fn(*self.static_inputs)

    def __call__(self, *inputs: torch.Tensor):
        for s, t in zip(self.static_inputs, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_outputs
