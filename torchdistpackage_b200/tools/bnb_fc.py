"""bitsandbytes int8 linear replacement (reference: tools/bnb_fc.py:1-22).  Optional: the library
is imported on first use (it is not installed in the B200 image; ``tools.int8_linear`` is the
self-contained alternative)."""
import torch.nn as nn

from .module_replace import replace_all_module


def _to_bnb(fc: nn.Linear) -> nn.Module:
    try:
        import bitsandbytes as bnb
    except ImportError as e:       # not in the B200 image: tools.int8_linear is the in-tree option
        raise ImportError("bitsandbytes is not installed; use tools.replace_linear_by_int8") from e
    has_bias = fc.bias is not None
    new = bnb.nn.Linear8bitLt(fc.in_features, fc.out_features, bias=has_bias,
                              has_fp16_weights=False, threshold=6.0)
    new.load_state_dict(fc.state_dict())
    return new.to(fc.weight.device)     # (the reference passes the dtype here by mistake)


def replace_linear_by_bnb(model: nn.Module) -> nn.Module:
    return replace_all_module(model, lambda m: isinstance(m, nn.Linear), _to_bnb)
