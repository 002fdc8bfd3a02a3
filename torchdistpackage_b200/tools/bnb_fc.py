"""bitsandbytes int8 linear replacement (reference: tools/bnb_fc.py:1-22).  Optional: the library
is imported on first use (it is not installed in the B200 image; ``tools.int8_linear`` is the
self-contained alternative)."""
import torch.nn as nn

from .module_replace import replace_all_module


_BNB_DEFAULTS = {"has_fp16_weights": False, "threshold": 6.0}


def if_replace_linear(module: nn.Module) -> bool:
    return isinstance(module, nn.Linear)


def _to_bnb(fc: nn.Linear, bnb_kwargs=None) -> nn.Module:
    try:
        import bitsandbytes as bnb
    except ImportError as e:       # not in the B200 image: tools.int8_linear is the in-tree option
        raise ImportError("bitsandbytes is not installed; use tools.replace_linear_by_int8") from e
    has_bias = fc.bias is not None
    new = bnb.nn.Linear8bitLt(fc.in_features, fc.out_features, bias=has_bias,
                              **(_BNB_DEFAULTS if bnb_kwargs is None else bnb_kwargs))
    new.load_state_dict(fc.state_dict())
    return new.to(fc.weight.device)     # (the reference passes the dtype here by mistake)


get_new_module = _to_bnb


def replace_linear_by_bnb(model: nn.Module, bnb_kwargs=None) -> nn.Module:
    """``bnb_kwargs`` go to ``bitsandbytes.nn.Linear8bitLt`` (default: int8 weights, outlier
    threshold 6.0 -- the reference's defaults, tools/bnb_fc.py:21)."""
    return replace_all_module(model, if_replace_linear, lambda m: _to_bnb(m, bnb_kwargs))
