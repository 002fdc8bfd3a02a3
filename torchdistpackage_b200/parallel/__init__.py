from .pipeline_parallel.pipeline_sched import forward_backward, forward_eval
from .pipeline_parallel.pipeline_helper import partition_uniform, partition_balanced, flatten_model
from .pipeline_parallel.clip_grad_parallel import clip_grad_norm_, NativeScalerPP

from .tensor_parallel.transformer import ParallelBlock, Block, Transformer
from .tensor_parallel.attn import Attention, TpAttention
from .tensor_parallel.mlp import Mlp, TpMlp
from .tensor_parallel.tp_utils import *  # noqa: F401,F403
