from .pipeline_sched import forward_backward, forward_eval
from .pipeline_helper import (partition_uniform, partition_balanced, flatten_model, flatten_sequence,
                              flat_and_partition, CallableModule, uniform_bounds, balanced_bounds)
from .clip_grad_parallel import clip_grad_norm_, NativeScalerPP
from . import comm
