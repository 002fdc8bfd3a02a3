"""Locate / load the native sm_100a extension (``torchdistpackage_b200/_C.so``).

Policy: on a machine with a GPU the native kernels ARE the product -- if the extension cannot be
imported we try one in-tree build and otherwise fail loudly (no silent PyTorch fallback).  On a
CPU-only machine (unit tests, gloo) ``native()`` returns ``None`` and callers take the reference
PyTorch path.
"""
from __future__ import annotations

import os
import threading

import torch

_lock = threading.Lock()
_mod = None
_tried = False


def _import():
    from .. import _C  # type: ignore
    return _C


def native(required: bool = False):
    """Return the extension module, or ``None`` if unavailable (and not ``required``)."""
    global _mod, _tried
    if os.environ.get("TDP_DISABLE_NATIVE") == "1" and not required:
        # explicit opt-out (used by baseline harnesses that must run on plain torch kernels)
        return None
    if _mod is not None:
        return _mod
    with _lock:
        if _mod is not None:
            return _mod
        if not _tried:
            _tried = True
            try:
                _mod = _import()
            except Exception as first:  # not built yet (fresh checkout) -> build in tree once
                if os.environ.get("TDP_NO_AUTOBUILD") != "1" and (
                        torch.cuda.is_available() or required):
                    try:
                        from ._build import build
                        build()
                        _mod = _import()
                    except Exception as second:
                        if torch.cuda.is_available() or required:
                            raise RuntimeError(
                                "torchdistpackage_b200: native sm_100a extension missing and the "
                                f"in-tree build failed.\nimport error: {first}\nbuild error: {second}"
                            ) from second
        if _mod is None and (required or torch.cuda.is_available()):
            raise RuntimeError("torchdistpackage_b200: native extension (_C.so) is not available; "
                               "run `python -m torchdistpackage_b200.ops._build`")
        return _mod


def have_native() -> bool:
    try:
        return native() is not None
    except Exception:
        return False


def use_native_for(t: torch.Tensor) -> bool:
    """Native kernels apply to CUDA tensors only; CPU tensors take the PyTorch path."""
    return t.is_cuda and native() is not None
