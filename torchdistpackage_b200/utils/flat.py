"""Flat-buffer helpers shared by the DDP buckets, ZeRO shards and the fused optimizer."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import torch


def align_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


@dataclass
class FlatView:
    """A tensor living at ``[offset, offset + numel)`` of a flat buffer."""
    name: str
    offset: int
    numel: int
    shape: torch.Size


def flatten_like(tensors: Sequence[torch.Tensor], dtype=None, device=None, align_elems: int = 1,
                 out: torch.Tensor = None):
    """Lay ``tensors`` out back to back (each start aligned to ``align_elems``) in one flat buffer.
    Returns ``(flat, views)`` where ``views[i]`` is a view of ``flat`` shaped like ``tensors[i]``
    holding a copy of its data."""
    offs, total = [], 0
    for t in tensors:
        total = align_up(total, align_elems)
        offs.append(total)
        total += t.numel()
    total = align_up(max(total, 1), align_elems)
    if out is None:
        dtype = dtype or tensors[0].dtype
        device = device or tensors[0].device
        out = torch.zeros(total, dtype=dtype, device=device)
    else:
        assert out.numel() >= total
    views: List[torch.Tensor] = []
    for t, o in zip(tensors, offs):
        v = out[o:o + t.numel()].view(t.shape)
        v.copy_(t.detach())
        views.append(v)
    return out, views
