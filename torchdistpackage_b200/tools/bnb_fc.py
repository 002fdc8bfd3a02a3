"""bitsandbytes int8 linear replacement (reference: tools/bnb_fc.py:1-22).  Optional: importing
this module fails cleanly when bitsandbytes is not installed (it is not part of the B200 image;
note that sm_100 has INT8 tensor cores but the bf16 / fp8 tcgen05 paths are what this package
targets)."""
import torch
import torch.nn as nn
import bitsandbytes as bnb  # noqa: F401  (ImportError is handled by the package root)

from .module_replace import replace_all_module


def _to_bnb(fc: nn.Linear) -> nn.Module:
    has_bias = fc.bias is not None
    new = bnb.nn.Linear8bitLt(fc.in_features, fc.out_features, bias=has_bias,
                              has_fp16_weights=False, threshold=6.0)
    new.load_state_dict(fc.state_dict())
    return new.to(fc.weight.device)     # (the reference passes the dtype here by mistake)


def replace_linear_by_bnb(model: nn.Module) -> nn.Module:
    return replace_all_module(model, lambda m: isinstance(m, nn.Linear), _to_bnb)
