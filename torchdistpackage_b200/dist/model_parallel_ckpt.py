"""Checkpoint helpers for model-parallel runs.

``get_mp_ckpt_suffix`` names per-(tp, pp) shards like the reference intended
(dist/model_parallel_ckpt.py:4-21; broken there: it calls an undefined ``is_mode_inited``).
On top of the naming helper this module provides a small save / load driver for sharded state:
every model-parallel rank of the *first* data-parallel replica writes its own file, ZeRO / EMA
shards add their data-parallel rank.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist

from .process_topo import tpc


def get_mp_ckpt_suffix(include_dp: bool = False) -> str:
    """``"_tp_{r}_pp_{r}.pth"`` (axes that are not in use are omitted)."""
    name = ""
    if tpc.is_mode_inited("tensor"):
        name += f"_tp_{tpc.get_group_rank('tensor')}"
    if tpc.is_mode_inited("pipe"):
        name += f"_pp_{tpc.get_group_rank('pipe')}"
    if include_dp and tpc.is_mode_inited("data"):
        name += f"_dp_{tpc.get_group_rank('data')}"
    return name + ".pth"


def _is_dp_writer() -> bool:
    return (not tpc.is_mode_inited("data")) or tpc.get_group_rank("data") == 0


def save_mp_checkpoint(prefix: str, model_state: Dict[str, Any],
                       sharded_state: Optional[Dict[str, Any]] = None) -> str:
    """Write ``{prefix}_tp_x_pp_y.pth`` from DP replica 0 and, if given, the per-DP-rank sharded
    state (ZeRO master weights / optimizer moments / EMA shard) to ``..._dp_z.pth``."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)) or ".", exist_ok=True)
    path = prefix + get_mp_ckpt_suffix()
    if _is_dp_writer():
        torch.save(model_state, path)
    if sharded_state is not None:
        torch.save(sharded_state, prefix + "_shard" + get_mp_ckpt_suffix(include_dp=True))
    if dist.is_initialized():
        dist.barrier()
    return path


def load_mp_checkpoint(prefix: str, map_location="cpu", with_shard: bool = False):
    state = torch.load(prefix + get_mp_ckpt_suffix(), map_location=map_location, weights_only=False)
    if not with_shard:
        return state
    shard = torch.load(prefix + "_shard" + get_mp_ckpt_suffix(include_dp=True),
                       map_location=map_location, weights_only=False)
    return state, shard


class AsyncCheckpointWriter:
    """Checkpoint without stalling the training loop.

    ``save(prefix, model_state, sharded_state)`` copies every tensor to (pinned, when CUDA is
    present) host memory on a side stream, records an event and returns; a background thread waits
    for the event and does the ``torch.save`` file I/O with the same naming / writer rules as
    :func:`save_mp_checkpoint`.  One save is in flight at a time (a second ``save`` first waits for
    the previous one), ``wait()`` blocks until the files are on disk, ``last_error`` keeps a
    failure of the background write.  Files are written to ``<name>.tmp`` and renamed, so a job
    killed mid-write never leaves a truncated checkpoint behind."""

    def __init__(self):
        import threading
        self._thread: Optional["threading.Thread"] = None
        self._stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.last_error: Optional[BaseException] = None
        self.last_paths: list = []

    def _to_host(self, obj):
        if isinstance(obj, torch.Tensor):
            if obj.is_cuda:
                host = torch.empty(obj.shape, dtype=obj.dtype, device="cpu", pin_memory=True)
                host.copy_(obj, non_blocking=True)
                return host
            return obj.detach().clone()
        if isinstance(obj, dict):
            return type(obj)((k, self._to_host(v)) for k, v in obj.items())
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._to_host(v) for v in obj)
        return obj

    def save(self, prefix: str, model_state: Dict[str, Any],
             sharded_state: Optional[Dict[str, Any]] = None) -> None:
        import threading
        self.wait()
        os.makedirs(os.path.dirname(os.path.abspath(prefix)) or ".", exist_ok=True)
        jobs = []
        event = None
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            ctx = torch.cuda.stream(self._stream)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            if _is_dp_writer():
                jobs.append((prefix + get_mp_ckpt_suffix(), self._to_host(model_state)))
            if sharded_state is not None:
                jobs.append((prefix + "_shard" + get_mp_ckpt_suffix(include_dp=True),
                             self._to_host(sharded_state)))
            if self._stream is not None:
                event = torch.cuda.Event()
                event.record(self._stream)

        def write():
            try:
                if event is not None:
                    event.synchronize()
                for path, state in jobs:
                    torch.save(state, path + ".tmp")
                    os.replace(path + ".tmp", path)
                self.last_paths = [p for p, _ in jobs]
            except BaseException as e:      # surfaced by wait()
                self.last_error = e

        self._thread = threading.Thread(target=write, name="tdp-async-ckpt", daemon=True)
        self._thread.start()

    def wait(self) -> None:
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        if self.last_error is not None:
            err, self.last_error = self.last_error, None
            raise RuntimeError("asynchronous checkpoint write failed") from err
