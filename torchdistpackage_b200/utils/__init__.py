"""General utilities: deterministic seeding and greedy parameter partitioning.

Parity: reference ``torchdistpackage/utils.py`` (``fix_rand`` :4-33, ``partition_params`` :35-64).
"""
from __future__ import annotations

import random
from typing import Dict, Iterable, List, Tuple, Union

import torch

from .flat import FlatView, flatten_like, align_up  # noqa: F401


def fix_rand(rank: int = 0, deterministic_cudnn: bool = True) -> int:
    """Seed python / numpy / torch (+cuda) with ``2222 + rank`` and make cuDNN deterministic."""
    seed = 2222 + int(rank)
    random.seed(seed)
    try:
        import numpy as np
        np.random.seed(seed)
    except Exception:  # numpy is optional
        pass
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if deterministic_cudnn:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
    return seed


def greedy_partition_sizes(numels: List[int], num_partitions: int) -> List[int]:
    """Assign consecutive items to ``num_partitions`` contiguous shards, closing a shard once it
    holds at least ``total / num_partitions`` elements.  Returns the owner index of each item.
    (Whole tensors are never split; this is the ZeRO / EMA sharding rule of the reference:
    ddp/zero_optim.py:19-41, utils.py:35-64.)"""
    total = sum(numels)
    target = total / max(num_partitions, 1)
    owners, cur, acc = [], 0, 0
    for n in numels:
        owners.append(cur)
        acc += n
        if acc >= target and cur < num_partitions - 1:
            cur += 1
            acc = 0
    return owners


def partition_params(model: Union[torch.nn.Module, Iterable[Tuple[str, torch.Tensor]]],
                     num_partitions: int, return_dict: bool = False):
    """Greedy numel-balanced partition of ``named_parameters`` into ``num_partitions`` shards.

    Returns a list (length ``num_partitions``) of lists of parameters, or of ``{name: param}``
    dicts when ``return_dict`` is set."""
    named = list(model.named_parameters()) if isinstance(model, torch.nn.Module) else list(model)
    owners = greedy_partition_sizes([p.numel() for _, p in named], num_partitions)
    parts: List = [dict() if return_dict else list() for _ in range(num_partitions)]
    for (name, p), o in zip(named, owners):
        if return_dict:
            parts[o][name] = p
        else:
            parts[o].append(p)
    return parts
