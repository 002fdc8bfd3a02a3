import argparse, atexit, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchdistpackage_b200 as tdp


def init(desc=""):
    ap = argparse.ArgumentParser(description=desc)
    ap.add_argument("--cpu", action="store_true")
    args, _ = ap.parse_known_args()
    rank, world, _, _ = tdp.setup_distributed("gloo" if args.cpu or not torch.cuda.is_available() else "nccl")
    dev = torch.device("cpu") if args.cpu or not torch.cuda.is_available() else \
        torch.device("cuda", torch.cuda.current_device())
    atexit.register(_shutdown)
    return rank, world, dev


def _shutdown():
    """Leave together: a rank that exits while a peer is still inside its last collective tears
    the peer's connection down (gloo aborts the process on a reset connection)."""
    import torch.distributed as dist
    if dist.is_initialized():
        try:
            dist.barrier()
        except Exception:
            pass
        tdp.shutdown_distributed()


def log(rank, *a):
    if rank == 0:
        print(*a, flush=True)
