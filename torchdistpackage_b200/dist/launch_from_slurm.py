"""Import-path compatibility: the reference keeps its bootstrap in
``torchdistpackage/dist/launch_from_slurm.py`` (:1-64); here it lives in :mod:`.launch` (SLURM,
torchrun and single-process start-up behind one function).  ``from
torchdistpackage_b200.dist.launch_from_slurm import setup_distributed`` keeps working."""
from .launch import find_free_port, setup_distributed, get_cpu_group, shutdown_distributed  # noqa: F401

__all__ = ["setup_distributed", "find_free_port", "get_cpu_group", "shutdown_distributed"]
