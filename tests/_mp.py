"""Spawn helper for multi-process tests: gloo on CPU (default) or nccl on GPUs."""
import os
import socket
import traceback

import torch
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, backend, fn, args, err_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    for k in list(os.environ):
        if k.startswith("SLURM_"):
            os.environ.pop(k)
    import _cov
    _cov.start()
    try:
        import torchdistpackage_b200 as tdp
        tdp.tpc.reset()
        tdp.tpc.verbose = False
        tdp.setup_distributed(backend)
        fn(rank, world, *args)
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        _cov.dump()
    except Exception:
        err_q.put((rank, traceback.format_exc()))
        raise


def run_distributed(fn, world: int, *args, backend: str = "gloo", timeout: float = 240.0):
    """Run ``fn(rank, world, *args)`` in ``world`` processes; raise if any rank failed."""
    ctx = mp.get_context("spawn")
    err_q = ctx.SimpleQueue()
    port = _free_port()
    procs = []
    for r in range(world):
        p = ctx.Process(target=_entry, args=(r, world, port, backend, fn, args, err_q))
        p.start()
        procs.append(p)
    failed = []
    for p in procs:
        p.join(timeout)
        if p.is_alive():
            p.terminate()
            failed.append("timeout")
        elif p.exitcode != 0:
            failed.append(f"exit {p.exitcode}")
    msgs = []
    while not err_q.empty():
        msgs.append(err_q.get())
    if failed or msgs:
        raise AssertionError("distributed test failed: " + "; ".join(failed) + "\n" +
                             "\n".join(f"[rank {r}]\n{m}" for r, m in msgs))
