"""Profiling / NVTX / printing helpers (reference: dist/utils.py:1-103).

``cu_prof_start/stop`` bracket a region for ``nsys --capture-range=cudaProfilerApi`` (or ncu),
``nvtx_decorator`` / ``NVTXContext`` push NVTX ranges (optionally with device-synchronised wall
time), ``_has_inf_or_nan`` is the overflow check, ``disable_non_master_print`` silences non-master
ranks (``print(..., force=True)`` still prints)."""
from __future__ import annotations

import builtins
import functools
import time

import torch

_builtin_print = builtins.print


def cu_prof_start() -> None:
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()


def cu_prof_stop() -> None:
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()


def _nvtx_push(name: str) -> None:
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)


def _nvtx_pop() -> None:
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_pop()


def nvtx_decorator(func=None, name: str = None):
    """NVTX range around every call of the decorated function.  Usable bare, as in the reference
    (``@nvtx_decorator``, dist/utils.py:35-45: the range is named after the function), or with a
    label (``@nvtx_decorator("fwd")`` / ``@nvtx_decorator(name="fwd")``)."""
    if isinstance(func, str):
        func, name = None, func

    def deco(fn):
        label = name or fn.__qualname__

        @functools.wraps(fn)
        def wrapped(*a, **k):
            _nvtx_push(label)
            try:
                return fn(*a, **k)
            finally:
                _nvtx_pop()
        return wrapped
    return deco(func) if callable(func) else deco


class NVTXContext:
    """``with NVTXContext('fwd', record_time=True) as c: ...`` -> ``c.elapsed_ms``"""

    def __init__(self, context_name: str, record_time: bool = False, enabled: bool = True):
        self.name = self.context_name = context_name
        self.record_time, self.enabled = record_time, enabled
        self.elapsed_ms = None

    def __enter__(self):
        if self.enabled:
            if self.record_time:
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                self._t0 = time.perf_counter()
            _nvtx_push(self.name)
        return self

    def __exit__(self, *exc):
        if self.enabled:
            _nvtx_pop()
            if self.record_time:
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                self.elapsed_ms = (time.perf_counter() - self._t0) * 1e3
                _builtin_print(f"[{self.name}] {self.elapsed_ms:.3f} ms", flush=True)
        return False


def _has_inf_or_nan(x: torch.Tensor) -> bool:
    try:
        s = float(x.float().sum())
    except RuntimeError as e:       # pragma: no cover
        if "value cannot be converted" not in str(e):
            raise
        return True
    return s in (float("inf"), -float("inf")) or s != s


def disable_non_master_print(is_master: bool) -> None:
    def _print(*args, **kwargs):
        force = kwargs.pop("force", False)
        if is_master or force:
            _builtin_print(*args, **kwargs)
    builtins.print = _print


def restore_print() -> None:
    builtins.print = _builtin_print


def report_memory(tag: str = "") -> dict:
    """Allocator snapshot in MiB (the helper the reference only had inside an example)."""
    if not torch.cuda.is_available():
        return {}
    d = dict(allocated=torch.cuda.memory_allocated() / 2 ** 20,
             max_allocated=torch.cuda.max_memory_allocated() / 2 ** 20,
             reserved=torch.cuda.memory_reserved() / 2 ** 20)
    _builtin_print(f"[mem {tag}] " + ", ".join(f"{k}={v:.0f}MiB" for k, v in d.items()), flush=True)
    return d
