"""Helpers that turn a model into a list of layers and hand each pipeline stage its slice.

Parity with the reference (parallel/pipeline_parallel/pipeline_helper.py): ``partition_uniform``,
``partition_balanced`` (by parameter count; the reference version calls an undefined
``_binary_partition``, :40), ``flatten_sequence``, ``CallableModule``, ``flatten_model``,
``flat_and_partition``.  The split math is exposed as pure functions
(``uniform_bounds`` / ``balanced_bounds``) so it can be tested without a process group.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.nn as nn

from ...dist.process_topo import tpc


# ---------------------------------------------------------------------------- pure split math
def uniform_bounds(n_items: int, n_parts: int, extra_len: int = 0) -> List[Tuple[int, int]]:
    """``(n_items + extra_len) // n_parts`` items per stage; the last stage takes the remainder
    (``extra_len`` lets callers reserve room, e.g. for an embedding that counts as a layer)."""
    per = (n_items + extra_len) // n_parts
    bounds = []
    for r in range(n_parts):
        beg = min(r * per, n_items)
        end = min((r + 1) * per, n_items) if r != n_parts - 1 else n_items
        bounds.append((beg, max(beg, end)))
    return bounds


def balanced_bounds(weights: Sequence[int], n_parts: int) -> List[Tuple[int, int]]:
    """Contiguous partition of ``weights`` into ``n_parts`` non-empty intervals minimising the
    heaviest interval (binary search on the bottleneck, then greedy fill, then split the
    heaviest multi-item intervals if fewer than ``n_parts`` were needed)."""
    n = len(weights)
    assert n >= n_parts > 0
    w = [max(int(x), 1) for x in weights]

    def parts_needed(cap: int):
        cuts, acc = [0], 0
        for i, x in enumerate(w):
            if acc + x > cap and acc > 0:
                cuts.append(i)
                acc = 0
            acc += x
        cuts.append(n)
        return cuts

    lo, hi = max(w), sum(w)
    while lo < hi:
        mid = (lo + hi) // 2
        if len(parts_needed(mid)) - 1 <= n_parts:
            hi = mid
        else:
            lo = mid + 1
    cuts = parts_needed(lo)
    intervals = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]
    while len(intervals) < n_parts:
        # split the heaviest interval that still has more than one item, as evenly as possible
        cand = [(sum(w[a:b]), idx) for idx, (a, b) in enumerate(intervals) if b - a > 1]
        _, idx = max(cand)
        a, b = intervals[idx]
        half, acc, cut = sum(w[a:b]) / 2, 0, a + 1
        for i in range(a, b - 1):
            acc += w[i]
            cut = i + 1
            if acc >= half:
                break
        intervals[idx:idx + 1] = [(a, cut), (cut, b)]
    return intervals


# ---------------------------------------------------------------------------- stage selection
def partition_uniform(flat_sequence: list, extra_len: int = 0) -> list:
    rank, world = tpc.get_group_rank("pipe"), tpc.get_group_size("pipe")
    beg, end = uniform_bounds(len(flat_sequence), world, extra_len)[rank]
    return flat_sequence[beg:end]


def partition_balanced(flat_sequence: list, sequence=None, **kwargs) -> list:
    rank, world = tpc.get_group_rank("pipe"), tpc.get_group_size("pipe")

    def n_params(m) -> int:
        return sum(p.numel() for p in m.parameters()) if isinstance(m, nn.Module) else 0

    beg, end = balanced_bounds([n_params(m) for m in flat_sequence], world)[rank]
    return flat_sequence[beg:end]


def flatten_sequence(sequence, level: int = 1) -> list:
    """Flatten nested ``nn.Sequential`` / lists ``level`` levels deep."""
    if level == 0:
        if isinstance(sequence, (list, tuple)):
            return list(sequence)
        if isinstance(sequence, (nn.Sequential, nn.ModuleList)):
            return list(sequence)
        return [sequence]
    if not isinstance(sequence, (list, tuple, nn.Sequential, nn.ModuleList)):
        return [sequence]
    out = []
    for el in sequence:
        out += flatten_sequence(el, level - 1)
    return out


class CallableModule(nn.Module):
    """Wrap a plain callable (lambda / function) so it can sit in an ``nn.Sequential``."""

    def __init__(self, fn: Callable):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


def flatten_model(model: nn.Module, layer_list: Sequence, return_list: bool = False):
    """Linearise a model given the execution order of its children, e.g. for a torchvision ResNet
    ``['conv1','bn1','relu','maxpool','layer1',...,'avgpool', lambda x: torch.flatten(x, 1), 'fc']``.
    Names resolve with ``getattr`` (containers are expanded), modules are used as is, callables
    are wrapped."""
    mods: List[nn.Module] = []
    for item in layer_list:
        if isinstance(item, str):
            sub = model.get_submodule(item) if "." in item else getattr(model, item)
            if isinstance(sub, (nn.Sequential, nn.ModuleList)):
                mods.extend(list(sub))
            else:
                mods.append(sub)
        elif isinstance(item, nn.Module):
            mods.append(item)
        elif callable(item):
            mods.append(CallableModule(item))
        else:
            raise NotImplementedError(f"flatten_model: unsupported entry {item!r}")
    return mods if return_list else nn.Sequential(*mods)


_POLICIES = {"uniform": partition_uniform, "balanced": partition_balanced}


def flat_and_partition(sequence, flat_level: int = 1, partition_policy: str = "uniform", **kwargs):
    flat = flatten_sequence(sequence, flat_level)
    if partition_policy not in _POLICIES:
        raise ValueError(f"unknown partition policy {partition_policy!r}")
    return _POLICIES[partition_policy](flat, **kwargs)
