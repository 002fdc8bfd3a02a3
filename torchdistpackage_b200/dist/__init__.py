from .launch import setup_distributed, find_free_port, get_cpu_group, shutdown_distributed
from .process_topo import (torch_parallel_context, tpc, ProcessTopology, test_comm, is_using_pp,
                           compute_layout, compute_axis_layout, compute_moe_layout)
from .node_group import (setup_node_groups, node_rank_lists, setup_inter_node_groups,
                         inter_node_rank_lists)
from .sharded_ema import ShardedEMA
from .model_parallel_ckpt import get_mp_ckpt_suffix, save_mp_checkpoint, load_mp_checkpoint
