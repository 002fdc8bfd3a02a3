"""NaN / Inf hunting hooks (reference: tools/debug_nan.py:1-60).  Instead of dropping into pdb
unconditionally the hooks raise ``FloatingPointError`` (set ``breakpoint_on_error=True`` to get
the reference's interactive behaviour)."""
from __future__ import annotations

import torch
import torch.nn as nn


def _bad(t: torch.Tensor) -> bool:
    return t.is_floating_point() and not bool(torch.isfinite(t).all())


def check_tensor_inf_nan(inp: torch.Tensor) -> bool:
    """True when ``t`` is clean -- no NaN / Inf (reference: tools/debug_nan.py:3-6)."""
    return not _bad(inp)


def check_tensors(inputs, where: str = "") -> bool:
    tensors = inputs
    """True when every tensor of the (nested) structure is finite, like the reference function of
    this name (:8-21, which looks one level deep; this one recurses through lists, tuples, dicts
    and objects carrying their payload in ``.sample``).  Offenders are reported with ``where``."""
    ok = True
    if isinstance(tensors, torch.Tensor):
        tensors = [tensors]
    elif isinstance(tensors, dict):
        tensors = list(tensors.values())
    elif hasattr(tensors, "sample") and not isinstance(tensors, (list, tuple)):
        tensors = [tensors.sample]
    for i, t in enumerate(tensors or []):
        if isinstance(t, torch.Tensor):
            if _bad(t):
                print(f"[debug_nan] non-finite values in {where or 'tensor'} (item {i}, shape "
                      f"{tuple(t.shape)})", flush=True)
                ok = False
        elif isinstance(t, (list, tuple, dict)) or hasattr(t, "sample"):
            ok = check_tensors(t, where) and ok
    return ok


def check_model_params(model: nn.Module) -> bool:
    """True when all parameters and their gradients are finite (the reference returns ``False``
    on the first bad parameter and falls off the end otherwise, :24-29)."""
    ok = True
    for n, p in model.named_parameters():
        if _bad(p.data):
            print(f"[debug_nan] parameter {n} is non-finite", flush=True)
            ok = False
        if p.grad is not None and _bad(p.grad):
            print(f"[debug_nan] gradient of {n} is non-finite", flush=True)
            ok = False
    return ok


def _fail(msg: str, breakpoint_on_error: bool):
    if breakpoint_on_error:
        import pdb
        pdb.set_trace()
    else:
        raise FloatingPointError(msg)


def fwd_hook_wrapper(module_name: str = "", breakpoint_on_error: bool = False):
    name = module_name

    def hook(module, inputs, output):
        if not (check_tensors(inputs, f"input of {name}")
                and check_tensors(output, f"output of {name}")):
            _fail(f"non-finite activation at {name}", breakpoint_on_error)
    return hook


def bwd_hook_wrapper(module_name: str = "", breakpoint_on_error: bool = False):
    name = module_name

    def hook(module, grad_input, grad_output):
        if not (check_tensors(grad_output, f"grad_output of {name}")
                and check_tensors(grad_input, f"grad_input of {name}")):
            _fail(f"non-finite gradient at {name}", breakpoint_on_error)
    return hook


def register_nan_hooks(model: nn.Module, breakpoint_on_error: bool = False):
    """Convenience: install both hooks on every sub-module; returns the handles."""
    hs = []
    for n, m in model.named_modules():
        hs.append(m.register_forward_hook(fwd_hook_wrapper(n, breakpoint_on_error)))
        hs.append(m.register_full_backward_hook(bwd_hook_wrapper(n, breakpoint_on_error)))
    return hs
