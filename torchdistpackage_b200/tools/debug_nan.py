"""NaN / Inf hunting hooks (reference: tools/debug_nan.py:1-60).  Instead of dropping into pdb
unconditionally the hooks raise ``FloatingPointError`` (set ``breakpoint_on_error=True`` to get
the reference's interactive behaviour)."""
from __future__ import annotations

import torch
import torch.nn as nn


def _bad(t: torch.Tensor) -> bool:
    return t.is_floating_point() and not bool(torch.isfinite(t).all())


def check_tensors(tensors, where: str = "") -> bool:
    """True if any tensor in the (nested) structure holds NaN/Inf; prints where."""
    found = False
    if isinstance(tensors, torch.Tensor):
        tensors = [tensors]
    if isinstance(tensors, dict):
        tensors = list(tensors.values())
    for i, t in enumerate(tensors or []):
        if isinstance(t, (list, tuple, dict)):
            found |= check_tensors(t, where)
        elif isinstance(t, torch.Tensor) and _bad(t):
            print(f"[debug_nan] non-finite values in {where} (item {i}, shape {tuple(t.shape)})",
                  flush=True)
            found = True
    return found


def check_model_params(model: nn.Module) -> bool:
    found = False
    for n, p in model.named_parameters():
        if _bad(p.data):
            print(f"[debug_nan] parameter {n} is non-finite", flush=True)
            found = True
        if p.grad is not None and _bad(p.grad):
            print(f"[debug_nan] gradient of {n} is non-finite", flush=True)
            found = True
    return found


def _fail(msg: str, breakpoint_on_error: bool):
    if breakpoint_on_error:
        import pdb
        pdb.set_trace()
    else:
        raise FloatingPointError(msg)


def fwd_hook_wrapper(name: str, breakpoint_on_error: bool = False):
    def hook(module, inputs, output):
        if check_tensors(inputs, f"input of {name}") or check_tensors(output, f"output of {name}"):
            _fail(f"non-finite activation at {name}", breakpoint_on_error)
    return hook


def bwd_hook_wrapper(name: str, breakpoint_on_error: bool = False):
    def hook(module, grad_input, grad_output):
        if check_tensors(grad_output, f"grad_output of {name}") or \
                check_tensors(grad_input, f"grad_input of {name}"):
            _fail(f"non-finite gradient at {name}", breakpoint_on_error)
    return hook


def register_nan_hooks(model: nn.Module, breakpoint_on_error: bool = False):
    """Convenience: install both hooks on every sub-module; returns the handles."""
    hs = []
    for n, m in model.named_modules():
        hs.append(m.register_forward_hook(fwd_hook_wrapper(n, breakpoint_on_error)))
        hs.append(m.register_full_backward_hook(bwd_hook_wrapper(n, breakpoint_on_error)))
    return hs
