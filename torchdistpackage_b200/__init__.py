"""torchdistpackage_b200 -- a Blackwell (B200, sm_100a) native mixed-parallel training toolkit with
the capabilities and API surface of KimmiShi/TorchDistPackage.

Public names mirror the reference package root (torchdistpackage/__init__.py:1-24) so that
``import torchdistpackage_b200 as torchdistpackage`` is a drop-in switch.
"""
from .ddp.naive_ddp import NaiveDDP, NaiveDdp, MoEDP, GradBucket, moe_dp_iter_step, create_moe_dp_hooks
from .ddp.zero_optim import Bf16ZeroOptimizer

from .dist.launch import setup_distributed, find_free_port, get_cpu_group, shutdown_distributed
from .dist.process_topo import torch_parallel_context as tpc
from .dist.process_topo import torch_parallel_context, test_comm, is_using_pp, ProcessTopology
from .dist.node_group import setup_node_groups, setup_inter_node_groups
from .dist.sharded_ema import ShardedEMA
from .dist.model_parallel_ckpt import get_mp_ckpt_suffix, save_mp_checkpoint, load_mp_checkpoint

from .utils import fix_rand, partition_params

from .tools.module_profiler import report_prof, register_profile_hooks, get_model_profile
from .tools.module_replace import replace_all_module
try:  # optional int8 back-ends (bitsandbytes / bminf are not part of this image)
    from .tools.bnb_fc import replace_linear_by_bnb
    from .tools.bminf_int8 import replace_linear_by_bminf
except Exception:  # pragma: no cover
    replace_linear_by_bnb = None
    replace_linear_by_bminf = None

from . import parallel  # noqa: E402
from . import ops  # noqa: E402

__version__ = "0.1.0"
