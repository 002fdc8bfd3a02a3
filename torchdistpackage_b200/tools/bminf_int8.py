"""BMInf int8 linear replacement (reference: tools/bminf_int8.py:1-14).  Optional import."""
import torch.nn as nn
import bminf  # noqa: F401  (ImportError is handled by the package root)

from .module_replace import replace_all_module


def replace_linear_by_bminf(model: nn.Module) -> nn.Module:
    return replace_all_module(model, lambda m: isinstance(m, nn.Linear),
                              lambda m: bminf.QuantizedLinear(m))
