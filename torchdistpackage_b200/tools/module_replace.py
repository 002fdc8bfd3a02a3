"""Recursive module surgery (reference: tools/module_replace.py:1-7)."""
from __future__ import annotations

from typing import Callable

import torch.nn as nn


def replace_all_module(model: nn.Module, if_replace_hook: Callable[[nn.Module], bool],
                       get_new_module: Callable[[nn.Module], nn.Module]) -> nn.Module:
    """Replace every sub-module for which ``if_replace_hook(m)`` is true by
    ``get_new_module(m)`` (children of a replaced module are not visited)."""
    for name, child in list(model.named_children()):
        if if_replace_hook(child):
            setattr(model, name, get_new_module(child))
        else:
            replace_all_module(child, if_replace_hook, get_new_module)
    return model
