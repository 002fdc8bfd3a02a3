"""Symmetric memory over NVSwitch: same-size buffers on every rank of a process group, each
rank holding peer-mapped pointers to all of them, one NVLS multicast mapping, and a signal pad
for in-kernel flags.  The collectives and the fused GEMM+collective kernels of this package
address these pointers directly (csrc/coll/collectives.cu, csrc/gemm/gemm_sm100.cuh).

Handle exchange (CUDA VMM export/import + multicast bind) is delegated to
``torch.distributed._symmetric_memory``; everything that moves data is our own kernel.

Signal pad word layout (uint32 words, see collectives.cu):
    [0, 1536)        barrier slots owned by the collective kernels
    [2048, ...)      user words: chunk flags / tile counters, handed out by ``alloc_words``
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from ._loader import native

SIGNAL_PAD_BYTES = 64 * 1024
USER_WORD_BASE = 4096        # first word after the barrier slots (3 x 128 blocks x 8 peers)
_MAX_WORLD = 8


def _dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.bfloat16:
        return 0
    if dtype == torch.float32:
        return 1
    raise TypeError(f"symmetric collectives support bf16 / fp32, got {dtype}")


_fd_serial = [0]


def _share_fd(pg, fd, owners):
    """Hand POSIX file descriptors to the other ranks of ``pg`` (same node) over abstract-namespace
    unix sockets with SCM_RIGHTS.  Every rank in ``owners`` serves its ``fd`` to all other ranks;
    returns ``{owner_rank: local duplicate}`` for the owners other than this rank.  (pidfd_getfd
    would be simpler but needs ptrace rights that containers usually drop.)"""
    import os
    import socket
    import threading
    rank, world = dist.get_rank(pg), dist.get_world_size(pg)
    _fd_serial[0] += 1
    addr = f"\0tdp-symm-{os.getpid()}-{_fd_serial[0]}"
    srv = None
    if rank in owners:
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(addr)
        srv.listen(world)
    addrs = [None] * world
    dist.all_gather_object(addrs, addr if srv is not None else None, group=pg)

    def serve():
        for _ in range(world - 1):
            conn, _ = srv.accept()
            with conn:
                socket.send_fds(conn, [b"f"], [fd])

    th = None
    if srv is not None:
        th = threading.Thread(target=serve, daemon=True)
        th.start()
    got = {}
    for r in owners:
        if r == rank:
            continue
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
            c.connect(addrs[r])
            _, fds, _, _ = socket.recv_fds(c, 16, 1)
            got[r] = fds[0]
    if th is not None:
        th.join()
        srv.close()
    return got


def _native_symm_alloc(pg, nbytes: int, want_multicast: bool):
    """Our own symmetric allocation (csrc/symm/symm_vmm.cpp): VMM allocation exported as a POSIX
    fd, fds passed between the ranks over unix sockets (SCM_RIGHTS), every peer's memory mapped
    here, one multicast object bound to all of them.  Layout: [data | 64 KiB signal pad].
    Returns ``(tensor, buffer_ptrs, signal_ptrs, multicast_ptr)``."""
    import os
    C = native(required=True)
    rank, world = dist.get_rank(pg), dist.get_world_size(pg)
    dev = torch.cuda.current_device()
    gran = C.vmm_granularity(dev, world)
    data = (nbytes + 4095) // 4096 * 4096
    total = (data + SIGNAL_PAD_BYTES + gran - 1) // gran * gran
    h, ptr, fd = C.vmm_alloc(total, dev)
    peer_fds = _share_fd(pg, fd, list(range(world)))
    ptrs = []
    for r in range(world):
        if r == rank:
            ptrs.append(ptr)
        else:
            _, pp = C.vmm_import(peer_fds[r], total, dev)
            os.close(peer_fds[r])
            ptrs.append(pp)
    mc_ptr = 0
    if want_multicast:
        ok, mch, mcfd = 1, None, -1
        if rank == 0:
            try:
                mch, mcfd = C.mc_create(total, world)
            except Exception:       # pragma: no cover - depends on driver / fabric support
                ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=pg)      # did rank 0 get its object?
        want_multicast = int(flag.item()) == 1
    if want_multicast:
        try:
            got = _share_fd(pg, mcfd, [0])
            if rank != 0:
                mch = C.mc_import(got[0])
                os.close(got[0])
            else:
                os.close(mcfd)
            C.mc_add_device(mch, dev)
        except Exception:           # pragma: no cover
            ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=pg)      # everyone added its device
        if int(flag.item()) == 1:
            mc_ptr = C.mc_bind_and_map(mch, h, total, dev)
    dist.barrier(group=pg)          # peers have imported: the exporter may close its fds
    os.close(fd)
    whole = C.tensor_from_ptr(ptr, total, dev)
    whole.zero_()
    tensor = whole[:nbytes]
    tensor._tdp_keepalive = whole
    return tensor, ptrs, [p + data for p in ptrs], mc_ptr


class SymmBuffer:
    """One symmetric allocation.  ``tensor`` is this rank's memory (uint8)."""

    def __init__(self, group: "SymmGroup", nbytes: int):
        import os
        self.group = group
        self.nbytes = int(nbytes)
        dev = torch.device("cuda", torch.cuda.current_device())
        backend = os.environ.get("TDP_SYMM_BACKEND", "torch")
        if backend == "native" and hasattr(native(required=True), "vmm_alloc"):
            self.tensor, bufs, sigs, mc = _native_symm_alloc(group.pg, self.nbytes,
                                                            not group.disable_multicast)
            self.handle = native(required=True).SymmHandle(bufs, sigs, mc, group.rank, group.world,
                                                           self.nbytes, dev.index)
            self.backend = "native"
        else:
            import torch.distributed._symmetric_memory as symm_mem
            self.tensor = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=dev)
            hdl = symm_mem.rendezvous(self.tensor, group.pg)
            self._torch_handle = hdl
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            if group.disable_multicast:
                mc = 0
            self.handle = native(required=True).SymmHandle(
                [int(p) for p in hdl.buffer_ptrs], [int(p) for p in hdl.signal_pad_ptrs], mc,
                int(hdl.rank), int(hdl.world_size), self.nbytes, dev.index)
            self.backend = "torch"
        self.has_multicast = bool(mc)
        self._next_word = USER_WORD_BASE
        self._epochs: Dict[int, int] = {}
        self.tensor.zero_()
        # all ranks must have finished initialising before anyone pushes data
        torch.cuda.synchronize()
        dist.barrier(group=group.pg)

    # ---------------------------------------------------------------- views
    def view(self, offset: int, shape, dtype: torch.dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        nb = n * torch.empty((), dtype=dtype).element_size()
        assert offset % 16 == 0 and offset + nb <= self.nbytes, "symmetric view out of range"
        return self.tensor[offset:offset + nb].view(dtype).view(*shape)

    def alloc_words(self, n: int) -> int:
        """Reserve ``n`` uint32 words of the signal pad (same index on every rank)."""
        w = self._next_word
        self._next_word += int(n)
        assert self._next_word * 4 <= SIGNAL_PAD_BYTES, "signal pad exhausted"
        return w

    def next_epoch(self, word: int, step: int = 1) -> int:
        """Monotonic per-flag epoch kept in lock step on all ranks by construction."""
        e = (self._epochs.get(word, 0) + step) & 0xFFFFFFFF
        self._epochs[word] = e
        return e

    # ---------------------------------------------------------------- collectives (in place)
    def barrier(self, slot: int = 0) -> None:
        self.handle.barrier(slot)

    def all_reduce_(self, offset: int, numel: int, dtype: torch.dtype, scale: float = 1.0,
                    algo: int = 0, max_ctas: int = 0) -> None:
        self.handle.all_reduce(offset, numel, _dtype_code(dtype), scale, algo, max_ctas)

    def reduce_scatter(self, offset: int, slice_numel: int, dtype: torch.dtype, out: torch.Tensor,
                       scale: float = 1.0, accumulate: bool = False, max_ctas: int = 0) -> None:
        self.handle.reduce_scatter(offset, slice_numel, _dtype_code(dtype), scale, out, accumulate,
                                   True, max_ctas)

    def all_gather(self, offset: int, slice_bytes: int, src: Optional[torch.Tensor] = None,
                   max_ctas: int = 0) -> None:
        self.handle.all_gather(offset, slice_bytes, src, True, max_ctas)


class SymmGroup:
    """Symmetric-memory context of one process group (NVSwitch domain, <= 8 ranks)."""

    def __init__(self, pg=None, disable_multicast: bool = False):
        self.pg = pg if pg is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.pg)
        self.world = dist.get_world_size(self.pg)
        self.disable_multicast = disable_multicast
        self.enabled = False
        self.reason = ""
        self._buffers: List[SymmBuffer] = []
        if not torch.cuda.is_available() or dist.get_backend(self.pg) != "nccl":
            self.reason = "no CUDA / not an NCCL group"
            return
        if self.world < 2 or self.world > _MAX_WORLD:
            self.reason = f"group size {self.world} outside [2, {_MAX_WORLD}]"
            return
        if native() is None:
            self.reason = "native extension missing"
            return
        # Peer mappings / NVLS multicast only exist inside one NVSwitch domain: every rank of
        # the group must sit on the same host (boot id) and see every peer device.  The check is
        # a collective, so all ranks agree and cross-node groups (e.g. the inter-node DDP group of
        # hybrid ZeRO, or DP across nodes under TP=8) fall back to NCCL instead of hanging in the
        # handle exchange.
        why = self._locality_problem()
        if why:
            self.reason = why
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem
            try:
                symm_mem.set_signal_pad_size(SIGNAL_PAD_BYTES)
            except Exception:
                pass
            try:
                symm_mem.enable_symm_mem_for_group(self.pg.group_name)
            except Exception:
                pass
            self.enabled = True
        except Exception as e:  # pragma: no cover - depends on the torch build
            self.reason = f"symmetric memory unavailable: {e}"

    def _locality_problem(self) -> str:
        host = _host_identity()
        dev = torch.cuda.current_device()
        try:
            uuid = str(torch.cuda.get_device_properties(dev).uuid)
        except Exception:
            uuid = f"{host}:{dev}"
        infos = [None] * self.world
        try:
            dist.all_gather_object(infos, (host, uuid), group=self.pg)
        except Exception as e:  # pragma: no cover
            return f"locality exchange failed: {e}"
        problem = ""
        if len({h for h, _ in infos}) != 1:
            problem = "group spans more than one host (no NVSwitch peer access)"
        elif len({u for _, u in infos}) != len(infos):
            problem = "two ranks of the group share a device"
        else:
            # peers this process can see (no CUDA_VISIBLE_DEVICES mask): ask the driver
            local = {}
            for i in range(torch.cuda.device_count()):
                try:
                    local[str(torch.cuda.get_device_properties(i).uuid)] = i
                except Exception:
                    pass
            for _, u in infos:
                d = local.get(u)
                if d is not None and d != dev and not torch.cuda.can_device_access_peer(dev, d):
                    problem = f"device {dev} cannot access peer device {d}"
                    break
        # agree on the outcome: one rank's failure disables the path everywhere
        flags = [None] * self.world
        dist.all_gather_object(flags, problem, group=self.pg)
        for f in flags:
            if f:
                return f
        return ""

    def alloc(self, nbytes: int) -> SymmBuffer:
        if not self.enabled:
            raise RuntimeError(f"symmetric memory not available for this group: {self.reason}")
        nbytes = (int(nbytes) + 2 * 1024 * 1024 - 1) // (2 * 1024 * 1024) * (2 * 1024 * 1024)
        buf = SymmBuffer(self, nbytes)
        self._buffers.append(buf)
        return buf


def _host_identity() -> str:
    """Stable per-boot identity of this host (containers on one box share it)."""
    try:
        with open("/proc/sys/kernel/random/boot_id") as f:
            return f.read().strip()
    except OSError:
        import socket
        return socket.gethostname()


_GROUP_CACHE: Dict[int, SymmGroup] = {}


def get_symm_group(pg=None) -> SymmGroup:
    """One :class:`SymmGroup` per process group (cached)."""
    key = id(pg if pg is not None else dist.group.WORLD)
    if key not in _GROUP_CACHE:
        _GROUP_CACHE[key] = SymmGroup(pg)
    return _GROUP_CACHE[key]


def reset_symm_cache() -> None:
    _GROUP_CACHE.clear()
