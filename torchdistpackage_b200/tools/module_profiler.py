"""Per-module forward time / activation-memory profiler.

Parity: reference ``register_profile_hooks`` / ``report_prof`` / ``get_model_profile``
(tools/module_profiler.py:1-171, tools/module_profile.md): forward pre/post hooks on every
sub-module measure wall time (device-synchronised) and ``memory_allocated`` growth minus the
output-vs-input activation delta; the report is grouped by hierarchy level and can be sorted by
``MB / ms`` to pick activation-checkpoint sites.

Differences: on GPU the times are CUDA-event pairs on the current stream, resolved once at
report time (the reference synchronises the device twice per module, which serialises the host
against every kernel and inflates small modules); backward times are collected as well;
``get_dt_size`` is correct for 1-byte dtypes (the reference returns 8 for int8, :23-24); works on
CPU (time only); hooks are removable.
"""
from __future__ import annotations

import time
from collections import OrderedDict, defaultdict
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn


class ProfileRecords(OrderedDict):
    """``{module name: record}`` -- what :func:`register_profile_hooks` returns (the reference
    returns its plain ``infos`` dict, :88-94).  A record holds ``fwd_time`` / ``bwd_time`` (ms,
    summed over calls), ``fwd_mem`` (MB), ``calls``, ``level`` and ``type``.  ``handles`` are the
    installed hooks; :meth:`remove` detaches them."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.handles: List[Any] = []
        self._pending: List[tuple] = []          # (record, key, start event, end event)

    def remove(self) -> None:
        for h in self.handles:
            h.remove()
        self.handles = []

    def resolve(self) -> None:
        """Turn recorded CUDA event pairs into milliseconds (one device synchronisation for the
        whole run instead of two per module)."""
        if self._pending:
            torch.cuda.synchronize()
            for rec, key, e0, e1 in self._pending:
                rec[key] += e0.elapsed_time(e1)
            self._pending = []


_RECORDS = ProfileRecords()          # the records of the most recent register_profile_hooks()
_ALL: List[ProfileRecords] = []


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


def count_tensor_size(output) -> int:
    """Bytes held by the tensors of a (nested) output structure (reference name, :27-36)."""
    return _tensor_bytes(output)


def output_same_as_input(output, args) -> bool:
    """A module that hands its input through (``nn.Identity``, in-place activations) adds no
    activation memory (reference helper, :38-45)."""
    if not isinstance(output, torch.Tensor):
        return False
    if output is args:
        return True
    return isinstance(args, (tuple, list)) and len(args) == 1 and output is args[0]


def get_level(name: str) -> int:
    """Hierarchy level of a dotted module name.  As in the reference (:52-57) an index into a
    ``ModuleList`` / ``Sequential`` does not open a new level -- ``blocks.3`` reports next to
    ``blocks`` and ``blocks.3.attn`` one level below -- but for indices of any width (the reference
    pattern only recognises single digits, so ``blocks.10.attn`` lands one level too deep)."""
    if name in ("", "root"):
        return 0
    parts = name.split(".")
    return 1 + sum(1 for c in parts[1:] if not c.isdigit())


def register_profile_hooks(model: nn.Module, infos: Optional[dict] = None,
                           max_depth: Optional[int] = None, backward: bool = False) -> ProfileRecords:
    """Install timing / memory hooks on ``model`` and all sub-modules (optionally only down to
    ``max_depth`` levels).  Returns the records (``infos`` itself when the caller passes its own
    :class:`ProfileRecords`, as the reference fills the dict it is given).  ``backward=True`` adds
    full-backward hooks for backward times (not for models with in-place activations: autograd
    rejects in-place edits of a hooked module's output)."""
    global _RECORDS
    recs = infos if isinstance(infos, ProfileRecords) else ProfileRecords()
    recs.clear()
    _RECORDS = recs
    _ALL.append(recs)
    cuda = torch.cuda.is_available() and any(p.is_cuda for p in model.parameters())

    def record(name, mod):
        key = name or "root"
        if key not in recs:
            recs[key] = dict(level=get_level(name), fwd_time=0.0, bwd_time=0.0, fwd_mem=0.0,
                             calls=0, type=type(mod).__name__)
        return recs[key]

    def begin(rec, key):
        if cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            rec["_" + key] = ev
        else:
            rec["_" + key] = time.perf_counter()

    def end(rec, key):
        t0 = rec.pop("_" + key, None)
        if t0 is None:
            return
        if cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            recs._pending.append((rec, key, t0, ev))
        else:
            rec[key] += (time.perf_counter() - t0) * 1e3

    def fwd_pre(name):
        def fn(mod, args):
            rec = record(name, mod)
            rec["_in_bytes"] = _tensor_bytes(args)
            if cuda:
                rec["_mem0"] = torch.cuda.memory_allocated()
            begin(rec, "fwd_time")
        return fn

    def fwd_post(name):
        def fn(mod, args, output):
            rec = record(name, mod)
            end(rec, "fwd_time")
            if cuda:
                grown = torch.cuda.memory_allocated() - rec.pop("_mem0", 0)
                act = 0 if output_same_as_input(output, args) else \
                    _tensor_bytes(output) - rec.pop("_in_bytes", 0)
                rec["fwd_mem"] += max(grown - max(act, 0), 0) / 1e6
            rec["calls"] += 1
        return fn

    def bwd_pre(name):
        def fn(mod, grad_output):
            begin(record(name, mod), "bwd_time")
        return fn

    def bwd_post(name):
        def fn(mod, grad_input, grad_output):
            end(record(name, mod), "bwd_time")
        return fn

    for name, mod in model.named_modules():
        if max_depth is not None and get_level(name) > max_depth:
            continue
        recs.handles.append(mod.register_forward_pre_hook(fwd_pre(name)))
        recs.handles.append(mod.register_forward_hook(fwd_post(name)))
        if backward:
            recs.handles.append(mod.register_full_backward_pre_hook(bwd_pre(name)))
            recs.handles.append(mod.register_full_backward_hook(bwd_post(name)))
    return recs


def remove_profile_hooks() -> None:
    """Detach the hooks of every :func:`register_profile_hooks` call so far."""
    for recs in _ALL:
        recs.remove()
    _ALL.clear()


def divide_by_layer(infos: Optional[Dict[str, dict]] = None) -> Dict[int, Dict[str, dict]]:
    """``{level: {name: record}}`` view of the records (reference name, :94-116)."""
    infos = _RECORDS if infos is None else infos
    out: Dict[int, Dict[str, dict]] = defaultdict(dict)
    for name, rec in infos.items():
        out[rec.get("level", get_level(name))][name] = rec
    return dict(sorted(out.items()))


def report_prof(infos: Optional[dict] = None, topn: Optional[int] = 20, max_depth: Optional[int] = 5,
                min_mem: float = 50, sort: bool = True, file=None) -> Dict[int, list]:
    """Print ``name  MB  fwd ms  bwd ms`` per hierarchy level (argument order and defaults of the
    reference's ``report_prof`` / ``sort_mem_time_ratio``, :118-142).  ``sort=True`` orders each
    level by MB per forward-ms -- the best activation-checkpoint candidates first -- keeps the
    ``topn`` best and drops modules holding less than ``min_mem`` MB; ``sort=False`` prints every
    module.  ``infos=None`` reports the most recent :func:`register_profile_hooks`.  Returns
    ``{level: [(name, mb, fwd_ms, bwd_ms), ...]}``."""
    infos = _RECORDS if infos is None else infos
    if isinstance(infos, ProfileRecords):
        infos.resolve()
    out: Dict[int, list] = {}
    for level, metas in divide_by_layer(infos).items():
        if max_depth is not None and level > max_depth:
            break
        rows = [(n, r["fwd_mem"], r["fwd_time"], r.get("bwd_time", 0.0)) for n, r in metas.items()]
        if sort:
            rows = [r for r in rows if r[1] >= min_mem]
            rows.sort(key=lambda r: r[1] / max(r[2], 1e-6), reverse=True)
            if topn:
                rows = rows[:topn]
        print(f"\nlevel: {level}", file=file)
        for name, mb, fms, bms in rows:
            print(f"{name}: MEM: {mb:.1f} MB; Time: {fms:.4f} ms fwd, {bms:.4f} ms bwd", file=file)
        out[level] = rows
    return out


sort_mem_time_ratio = report_prof      # the reference's other name for the report (:118,142)


def get_model_profile(model: nn.Module, args: tuple = (), kwargs: Optional[dict] = None,
                      sort: bool = True, topn: Optional[int] = 20, max_depth: Optional[int] = 5,
                      min_mem: float = 50, backward: bool = False):
    """One warm-up forward, then one forward of ``model(*args, **kwargs)`` under the profiler, and
    the report (reference: :144-171, same defaults).  ``backward=True`` also runs
    ``output.sum().backward()`` so the report carries backward times."""
    kwargs = kwargs or {}

    def run():
        with torch.set_grad_enabled(backward):
            out = model(*args, **kwargs)
            if backward:
                loss = out if isinstance(out, torch.Tensor) else out[0]
                loss.float().sum().backward()

    run()                          # warm-up, backward included (allocator, library heuristics:
    #                                the first backward GEMM of a shape costs milliseconds)
    if backward:
        model.zero_grad(set_to_none=True)
    recs = register_profile_hooks(model, max_depth=None, backward=backward)
    try:
        run()
    finally:
        recs.remove()
    return report_prof(recs, sort=sort, topn=topn, max_depth=max_depth, min_mem=min_mem)
