"""Collective bandwidth micro-benchmark: NCCL *and* this package's NVSwitch kernels.

Parity: reference ``test_collection`` / ``test_all2all_balanced`` (dist/py_comm_test.py:1-84):
fp16/bf16 payload, NCCL-tests style bus bandwidth ``busbw = algbw * frac * (n-1)/n`` with
``frac`` = 2 for all_reduce, 1 for all_gather / reduce_scatter / all_to_all.  The reference's
``reduce_scatter`` mode passes wrong arguments (:20-33); this one works.

Differences: timing is on the device (CUDA events, max over ranks, warm-up iterations) and each
mode is also measured through the symmetric-memory kernels (``impl='symm'``) when available.

    torchrun --nproc-per-node 8 -m torchdistpackage_b200.dist.py_comm_test --mode all_reduce --mib 256
"""
from __future__ import annotations

import argparse
import json
import time

import torch
import torch.distributed as dist

from .launch import setup_distributed

# bus bandwidth = algorithm bandwidth x fraction x (n-1)/n  (the NCCL-tests convention the
# reference uses, py_comm_test.py:10-17)
_FRAC = {"all_reduce": 2.0, "all_gather": 1.0, "reduce_scatter": 1.0, "all_to_all": 1.0}
mode_2_frac = _FRAC


def bus_bandwidth_gbs(mode: str, total_bytes: int, seconds: float, n: int) -> float:
    algbw = total_bytes / max(seconds, 1e-12) / 1e9
    return algbw * _FRAC[mode] * (n - 1) / max(n, 1)


class CommResult(tuple):
    """What :func:`test_collection` returns.  Unpacks like the reference's return value --
    ``bw, time_avg = test_collection(...)`` = (bus bandwidth in GB/s, seconds per call),
    py_comm_test.py:48-57 -- and carries the full record: ``res["ms"]``, ``res["algbw_gbs"]``,
    ``res["busbw_gbs"]``, ``res["bytes"]``, ``res["world"]``, ... or ``res.info`` as a dict."""

    def __new__(cls, info: dict):
        obj = super().__new__(cls, (round(info["busbw_gbs"], 3), info["ms"] * 1e-3))
        obj.info = dict(info)
        return obj

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.info[key]
        return super().__getitem__(key)

    def keys(self):
        return self.info.keys()


def _timed(fn, warmup: int, iters: int, device) -> float:
    for _ in range(warmup):
        fn()
    if device.type == "cuda":
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        sec = s.elapsed_time(e) / 1e3 / iters
    else:
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        sec = (time.perf_counter() - t0) / iters
    t = torch.tensor([sec], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def test_collection(ele_num_total: int, mode: str = "all_reduce", group=None, impl: str = "nccl",
                    dtype=torch.bfloat16, warmup: int = 3, iters: int = 10,
                    verbose: bool = True) -> CommResult:
    """Measure one collective.  ``ele_num_total`` = elements of the *full* (gathered / reduced)
    tensor.  Returns a :class:`CommResult` (``(busbw GB/s, seconds)`` + the full record)."""
    numel_total = int(ele_num_total)
    world = dist.get_world_size(group)
    cuda = torch.cuda.is_available() and dist.get_backend(group) == "nccl"
    device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    if not cuda and dtype == torch.bfloat16:
        dtype = torch.float32
    esize = torch.empty((), dtype=dtype).element_size()
    numel_total = numel_total // (world * 8) * (world * 8)
    part = numel_total // world
    full = torch.randn(numel_total, device=device).to(dtype)
    shard = torch.randn(part, device=device).to(dtype)

    if impl == "symm":
        from ..ops.symm import get_symm_group
        sg = get_symm_group(group)
        if not sg.enabled:
            raise RuntimeError(f"symmetric memory unavailable: {sg.reason}")
        buf = sg.alloc(numel_total * esize)
        buf.view(0, (numel_total,), dtype).copy_(full)
        out = torch.empty(part, dtype=dtype, device=device)
        fns = {
            "all_reduce": lambda: buf.all_reduce_(0, numel_total, dtype),
            "reduce_scatter": lambda: buf.reduce_scatter(0, part, dtype, out),
            "all_gather": lambda: buf.all_gather(0, part * esize, shard),
        }
        if mode not in fns:
            raise ValueError(f"symm impl does not provide {mode}")
        fn = fns[mode]
    else:
        out = torch.empty(part, dtype=dtype, device=device)
        if mode == "all_reduce":
            fn = lambda: dist.all_reduce(full, group=group)
        elif mode == "all_gather":
            if cuda:
                fn = lambda: dist.all_gather_into_tensor(full, shard, group=group)
            else:
                outs = list(full.chunk(world))
                fn = lambda: dist.all_gather(outs, shard, group=group)
        elif mode == "reduce_scatter":
            if cuda:
                fn = lambda: dist.reduce_scatter_tensor(out, full, group=group)
            else:
                fn = lambda: dist.all_reduce(full, group=group)   # gloo: emulate
        elif mode == "all_to_all":
            recv = torch.empty_like(full)
            fn = lambda: dist.all_to_all_single(recv, full, group=group)
        else:
            raise ValueError(mode)
    sec = _timed(fn, warmup, iters, device)
    total_bytes = numel_total * esize
    res = dict(mode=mode, impl=impl, world=world, bytes=total_bytes, ms=sec * 1e3,
               algbw_gbs=total_bytes / sec / 1e9,
               busbw_gbs=bus_bandwidth_gbs(mode, total_bytes, sec, world))
    if verbose and dist.get_rank() == 0:
        print(json.dumps(res), flush=True)
    return CommResult(res)


def test_all2all_balanced(ele_num: int, group=None, **kw):
    """Balanced ``all_to_all_single`` of ``ele_num`` elements per rank (reference: :60-78)."""
    return test_collection(ele_num, "all_to_all", group, **kw)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="all_reduce", choices=sorted(_FRAC))
    ap.add_argument("--mib", type=float, default=256.0, help="size of the full tensor in MiB")
    ap.add_argument("--impl", default="both", choices=["nccl", "symm", "both"])
    args = ap.parse_args()
    setup_distributed("auto")
    numel = int(args.mib * 2 ** 20 / 2)
    for impl in (["nccl", "symm"] if args.impl == "both" else [args.impl]):
        try:
            test_collection(numel, args.mode, impl=impl)
        except (RuntimeError, ValueError) as e:
            if dist.get_rank() == 0:
                print(f"[{impl}] skipped: {e}", flush=True)
    dist.barrier()


if __name__ == "__main__":
    main()
