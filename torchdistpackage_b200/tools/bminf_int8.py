"""BMInf int8 linear replacement (reference: tools/bminf_int8.py:1-14).  Optional: the library is
imported on first use; ``tools.int8_linear`` is the self-contained alternative."""
import torch.nn as nn

from .module_replace import replace_all_module


def _to_bminf(fc: nn.Linear) -> nn.Module:
    try:
        import bminf
    except ImportError as e:
        raise ImportError("bminf is not installed; use tools.replace_linear_by_int8") from e
    return bminf.QuantizedLinear(fc)


def replace_linear_by_bminf(model: nn.Module) -> nn.Module:
    return replace_all_module(model, lambda m: isinstance(m, nn.Linear), _to_bminf)
