"""Per-module forward time / activation-memory profiler.

Parity: reference ``register_profile_hooks`` / ``report_prof`` / ``get_model_profile``
(tools/module_profiler.py:1-171, tools/module_profile.md): forward pre/post hooks on every
sub-module measure wall time (device-synchronised) and ``memory_allocated`` growth minus the
output-vs-input activation delta; the report is grouped by hierarchy level and can be sorted by
``MB / ms`` to pick activation-checkpoint sites.

Differences: time is taken with CUDA events on the current stream when on GPU (one sync per
module instead of two), ``get_dt_size`` is correct for 1-byte dtypes (the reference returns 8 for
int8, :23-24), works on CPU (time only), hooks are removable.
"""
from __future__ import annotations

import time
from collections import OrderedDict, defaultdict
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

_RECORDS: "OrderedDict[str, dict]" = OrderedDict()
_HANDLES: List[Any] = []


def get_dt_size(dtype: torch.dtype) -> int:
    return torch.empty((), dtype=dtype).element_size()


def _tensor_bytes(obj) -> int:
    if isinstance(obj, torch.Tensor):
        return obj.numel() * obj.element_size()
    if isinstance(obj, (list, tuple)):
        return sum(_tensor_bytes(o) for o in obj)
    if isinstance(obj, dict):
        return sum(_tensor_bytes(o) for o in obj.values())
    return 0


def _level(name: str) -> int:
    return 0 if name == "" else name.count(".") + 1


def register_profile_hooks(model: nn.Module, max_depth: Optional[int] = None) -> List[Any]:
    """Install timing / memory hooks on ``model`` and all sub-modules (optionally only down to
    ``max_depth`` levels).  Returns the hook handles."""
    _RECORDS.clear()
    cuda = torch.cuda.is_available() and any(p.is_cuda for p in model.parameters())

    def pre(name):
        def fn(mod, args, kwargs=None):
            rec = _RECORDS.setdefault(name or "root", dict(level=_level(name), time_ms=0.0,
                                                           mem_mb=0.0, calls=0,
                                                           type=type(mod).__name__))
            rec["_in_bytes"] = _tensor_bytes(args)
            if cuda:
                torch.cuda.synchronize()
                rec["_mem0"] = torch.cuda.memory_allocated()
            rec["_t0"] = time.perf_counter()
        return fn

    def post(name):
        def fn(mod, args, output):
            rec = _RECORDS[name or "root"]
            if cuda:
                torch.cuda.synchronize()
            rec["time_ms"] += (time.perf_counter() - rec.pop("_t0")) * 1e3
            if cuda:
                grown = torch.cuda.memory_allocated() - rec.pop("_mem0")
                act_delta = _tensor_bytes(output) - rec.pop("_in_bytes", 0)
                rec["mem_mb"] += max(grown - max(act_delta, 0), 0) / 2 ** 20
            rec["calls"] += 1
        return fn

    handles = []
    for name, mod in model.named_modules():
        if max_depth is not None and _level(name) > max_depth:
            continue
        handles.append(mod.register_forward_pre_hook(pre(name)))
        handles.append(mod.register_forward_hook(post(name)))
    _HANDLES.extend(handles)
    return handles


def remove_profile_hooks() -> None:
    for h in _HANDLES:
        h.remove()
    _HANDLES.clear()


def report_prof(sort: bool = False, topn: Optional[int] = None, max_depth: Optional[int] = None,
                min_mem: float = 0.0, file=None) -> Dict[int, list]:
    """Print ``name  MB  ms`` per hierarchy level; ``sort=True`` orders each level by MB/ms (the
    best activation-checkpoint candidates first).  Returns ``{level: [(name, mb, ms), ...]}``."""
    by_level: Dict[int, list] = defaultdict(list)
    for name, rec in _RECORDS.items():
        if max_depth is not None and rec["level"] > max_depth:
            continue
        if rec["mem_mb"] < min_mem:
            continue
        by_level[rec["level"]].append((name, rec["mem_mb"], rec["time_ms"]))
    out = {}
    for level in sorted(by_level):
        rows = by_level[level]
        if sort:
            rows = sorted(rows, key=lambda r: r[1] / max(r[2], 1e-6), reverse=True)
        if topn:
            rows = rows[:topn]
        print(f"---- level {level} ----", file=file)
        for name, mb, ms in rows:
            print(f"{name:<48s} {mb:10.1f} MB {ms:10.4f} ms", file=file)
        out[level] = rows
    return out


def get_model_profile(model: nn.Module, args=(), kwargs=None, sort: bool = False,
                      topn: Optional[int] = None, max_depth: Optional[int] = None,
                      min_mem: float = 0.0):
    """Run one forward of ``model(*args, **kwargs)`` under the profiler and print the report."""
    kwargs = kwargs or {}
    handles = register_profile_hooks(model, max_depth=max_depth)
    try:
        with torch.no_grad():
            model(*args, **kwargs)
    finally:
        for h in handles:
            h.remove()
    return report_prof(sort=sort, topn=topn, max_depth=max_depth, min_mem=min_mem)
