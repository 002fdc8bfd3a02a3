"""Process bootstrap: one process per GPU, ``torch.distributed`` over NCCL (GPU) or gloo (CPU).

Capability parity with the reference's ``setup_distributed`` (dist/launch_from_slurm.py:16-62):
SLURM env (``SLURM_PROCID`` / ``SLURM_NTASKS`` / ``SLURM_NODELIST``) or torchrun env
(``RANK`` / ``WORLD_SIZE`` / ``MASTER_*``), returning ``(rank, world_size, port, addr)``.

Differences (deliberate):
* ``addr`` is always bound (the reference raises ``UnboundLocalError`` under torchrun, :62);
* ``backend="auto"`` picks nccl when CUDA is present and gloo otherwise, and with nccl a gloo
  side-group is available through :func:`get_cpu_group` for host-side metadata exchange
  (pipeline shape metadata, checkpoint gathers) without touching the GPU stream;
* a single-process mode (no env at all) initialises a world of one so every component can be
  exercised in unit tests;
* ``LOCAL_RANK`` is honoured for device selection before falling back to ``rank % n_gpus``.
"""
from __future__ import annotations

import os
import socket
import subprocess
from datetime import timedelta
from typing import Optional, Tuple

import torch
import torch.distributed as dist

DEFAULT_PORT = 54647  # same default as the reference (launch_from_slurm.py:43)

_CPU_GROUP = None


def find_free_port() -> int:
    """Ask the kernel for an unused TCP port (reference: launch_from_slurm.py:8-14)."""
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(("", 0))
        return int(s.getsockname()[1])


def _first_slurm_host(nodelist: str) -> str:
    try:
        out = subprocess.run(["scontrol", "show", "hostname", nodelist], check=True,
                             capture_output=True, text=True).stdout.split()
        if out:
            return out[0]
    except Exception:
        pass
    # fall back to parsing "node[01-04],other" by hand
    head = nodelist.split(",")[0]
    if "[" in head:
        prefix, rest = head.split("[", 1)
        first = rest.rstrip("]").split(",")[0].split("-")[0]
        return prefix + first
    return head


def _pick_backend(backend: str) -> str:
    if backend in (None, "auto"):
        return "nccl" if torch.cuda.is_available() else "gloo"
    return backend


def setup_distributed(backend: str = "nccl", port: Optional[int] = None,
                      timeout_s: int = 600) -> Tuple[int, int, int, str]:
    """Initialise the default process group and bind this process to its GPU.

    Returns ``(rank, world_size, port, addr)``.
    """
    backend = _pick_backend(backend)
    if backend == "nccl" and not torch.cuda.is_available():
        # CPU-only host (unit tests / BASELINE config #1): degrade instead of crashing
        backend = "gloo"

    if "SLURM_JOB_ID" in os.environ and "SLURM_PROCID" in os.environ and "RANK" not in os.environ:
        rank = int(os.environ["SLURM_PROCID"])
        world_size = int(os.environ["SLURM_NTASKS"])
        addr = os.environ.get("MASTER_ADDR") or _first_slurm_host(os.environ["SLURM_NODELIST"])
        if port is None:
            port = int(os.environ.get("MASTER_PORT", DEFAULT_PORT))
        local_rank = int(os.environ.get("SLURM_LOCALID", rank % max(torch.cuda.device_count(), 1)))
    elif "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        rank = int(os.environ["RANK"])
        world_size = int(os.environ["WORLD_SIZE"])
        addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
        if port is None:
            port = int(os.environ.get("MASTER_PORT", DEFAULT_PORT))
        local_rank = int(os.environ.get("LOCAL_RANK", rank % max(torch.cuda.device_count(), 1)))
    else:
        rank, world_size, addr, local_rank = 0, 1, "127.0.0.1", 0
        if port is None:
            port = find_free_port()

    os.environ["MASTER_ADDR"] = str(addr)
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    os.environ["WORLD_SIZE"] = str(world_size)
    os.environ.setdefault("LOCAL_RANK", str(local_rank))

    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank % torch.cuda.device_count())

    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world_size,
                                timeout=timedelta(seconds=timeout_s))
    return rank, world_size, int(port), str(addr)


def get_cpu_group():
    """A gloo group spanning the world, for host-side object / metadata exchange."""
    global _CPU_GROUP
    if not dist.is_initialized():
        return None
    if dist.get_backend() == "gloo":
        return dist.group.WORLD
    if _CPU_GROUP is None:
        _CPU_GROUP = dist.new_group(backend="gloo")
    return _CPU_GROUP


def shutdown_distributed() -> None:
    global _CPU_GROUP
    _CPU_GROUP = None
    if dist.is_initialized():
        dist.destroy_process_group()
