from .naive_ddp import NaiveDDP, NaiveDdp, MoEDP, GradBucket, create_moe_dp_hooks, moe_dp_iter_step
from .zero_optim import Bf16ZeroOptimizer
