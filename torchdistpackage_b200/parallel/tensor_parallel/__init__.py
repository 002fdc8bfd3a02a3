from .tp_utils import *  # noqa: F401,F403
from .tp_utils import (TpLinear, ColParallelLinear, RowParallelLinear, get_tp_group, set_tp_group)
from .attn import Attention, TpAttention
from .mlp import Mlp, TpMlp
from .transformer import Block, ParallelBlock, Transformer, allreduce_sequence_parallel_grads
