"""Intra-node process groups -- the enabler for *hybrid* ZeRO (shard optimizer state inside a
node over NVSwitch, plain data parallel across nodes).

Parity: reference ``setup_node_groups`` (dist/node_group.py:3-33) -- one group per consecutive
``num_per_node`` ranks; returns this rank's group, or ``None`` when the world does not span more
than one node (``world_size <= num_per_node``) or is not divisible.

Typical use::

    node_group = setup_node_groups()               # None on a single 8-GPU box
    zero_group = node_group or tpc.get_group('data')
    optim = Bf16ZeroOptimizer(torch.optim.AdamW(model.parameters()), dp_group=zero_group)
"""
from __future__ import annotations

from typing import List, Optional

import torch.distributed as dist


def node_rank_lists(world_size: int, num_per_node: int = 8) -> Optional[List[List[int]]]:
    """Pure helper: rank lists per node, or ``None`` if node groups make no sense."""
    if num_per_node <= 0 or world_size % num_per_node != 0 or world_size <= num_per_node:
        return None
    return [list(range(n * num_per_node, (n + 1) * num_per_node))
            for n in range(world_size // num_per_node)]


def setup_node_groups(num_per_node: int = 8):
    lists = node_rank_lists(dist.get_world_size(), num_per_node)
    if lists is None:
        return None
    mine = None
    for ranks in lists:
        grp = dist.new_group(ranks)  # collective over the whole world
        if dist.get_rank() in ranks:
            mine = grp
    return mine


def inter_node_rank_lists(world_size: int, num_per_node: int = 8) -> Optional[List[List[int]]]:
    """Pure helper: for every local index, the ranks holding it on each node."""
    if node_rank_lists(world_size, num_per_node) is None:
        return None
    return [list(range(i, world_size, num_per_node)) for i in range(num_per_node)]


def setup_inter_node_groups(num_per_node: int = 8):
    """The complementary axis of :func:`setup_node_groups`: one group per local index, joining
    the ranks that hold the same ZeRO shard on different nodes (``Bf16ZeroOptimizer(...,
    outer_group=...)`` or ``NaiveDDP(process_group=...)`` for hybrid ZeRO).  ``None`` on a
    single node."""
    lists = inter_node_rank_lists(dist.get_world_size(), num_per_node)
    if lists is None:
        return None
    mine = None
    for ranks in lists:
        grp = dist.new_group(ranks)
        if dist.get_rank() in ranks:
            mine = grp
    return mine
