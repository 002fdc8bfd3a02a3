"""Native sm_100a operator layer (see csrc/) and its Python front-ends."""
from ._loader import native, have_native
from . import linear, fused, symm
from .linear import gemm, linear as linear_fn, mlp as mlp_fn
from .fused import layer_norm, cross_entropy, FusedAdamW, flatten_module_params
from .symm import SymmGroup, SymmBuffer, get_symm_group
