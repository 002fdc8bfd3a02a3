"""Data parallelism: bucketed gradient all-reduce overlapped with backward.

API parity with the reference (ddp/naive_ddp.py): ``NaiveDDP(module, sync=False,
bucket_cap_mb=25, gradient_as_bucket_view=False, process_group=None, dp_rank0=0,
reduce_op="avg", verbose=False, num_grad_acc_iter=1)`` + ``reduce_gradients()`` /
``broadcast_params()``; ``MoEDP`` / ``create_moe_dp_hooks`` / ``moe_dp_iter_step`` for
replicated-expert data parallelism (ddp/moe_dp.md).

B200-first design (what is different from the reference's Python-over-NCCL engine):

* On GPU the buckets live in **NVSwitch symmetric memory** and a ready bucket is reduced by our
  own NVLS kernel (``multimem.ld_reduce`` in the switch + ``multimem.st`` broadcast, 1/N fused)
  on a high-priority side stream -- no NCCL on this path (csrc/coll/collectives.cu).
* Gradients are *born* in the bucket: ``p.grad`` is a view of the flat buffer from the first
  iteration on, and a post-accumulate hook re-attaches the view if the user (or
  ``zero_grad(set_to_none=True)``) dropped it -- the reference silently reduces stale memory in
  that case (naive_ddp.py:154,165-169).
* Only the bytes that are in use are reduced (the reference reduces the full 25 MiB capacity).
* ``reduce_gradients()`` orders the compute stream after the comm stream with an event; it does
  not ``cuda.synchronize()`` the device.
* CPU / gloo works (SUM then divide -- gloo has no AVG), which the reference cannot do.
* ``reduce_op="sum"`` really sums (reference bug: ``reduce_op.lower == "sum"`` is never true).
"""
from __future__ import annotations

import time
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops._loader import native
from ..ops.symm import get_symm_group
from ..utils.flat import align_up

_ALIGN_BYTES = 512  # bucket slot alignment (also satisfies the 16-byte vector kernels)


class GradBucket:
    """A flat gradient buffer holding several parameters' grads back to back.

    ``buffer`` may be a view into symmetric memory (``symm`` = (SymmBuffer, byte offset))."""

    def __init__(self, index: int, dtype: torch.dtype, device: torch.device, group,
                 capacity_numel: int, buffer: Optional[torch.Tensor] = None, symm=None):
        self.index = index
        self.dtype = dtype
        self.device = device
        self.group = group
        self.capacity = int(capacity_numel)
        self.buffer = buffer if buffer is not None else torch.zeros(self.capacity, dtype=dtype,
                                                                     device=device)
        self.symm = symm
        self.names: List[str] = []
        self.views: Dict[str, torch.Tensor] = {}
        self.used = 0
        self._ready = 0
        self.pending_work = None
        self.reduced = False
        # set by BucketAdamW(fused_comm=True): replaces the all-reduce of this bucket with the
        # fused reduce-scatter -> AdamW -> all-gather kernel
        self.fused_step = None
        # streams on which directly written gradients of this bucket were produced (ops.linear
        # note_grad_stream): the reduction is ordered behind each of them
        self.producer_streams = set()

    # ---- layout
    def get_aligned_size(self, tensor: torch.Tensor) -> int:
        """Elements ``tensor`` occupies in the bucket including the alignment padding behind it
        (reference helper of this name, naive_ddp.py:456-461)."""
        return align_up(tensor.numel(), self.elem_align())

    def elem_align(self) -> int:
        return max(1, _ALIGN_BYTES // self.buffer.element_size())

    def can_fit(self, numel: int) -> bool:
        return align_up(self.used, self.elem_align()) + numel <= self.capacity

    def push(self, name: str, shape: torch.Size, numel: int) -> torch.Tensor:
        start = align_up(self.used, self.elem_align())
        view = self.buffer[start:start + numel].view(shape)
        self.used = start + numel
        self.names.append(name)
        self.views[name] = view
        return view

    def payload(self) -> torch.Tensor:
        """The in-use prefix (what actually travels)."""
        n = align_up(self.used, max(self.elem_align(), 8))
        return self.buffer[:min(n, self.capacity)]

    # ---- readiness accounting
    def grad_ready(self) -> bool:
        self._ready += 1
        return self._ready >= len(self.names)

    def grad_reset(self) -> None:
        self._ready = 0
        self.reduced = False
        self.pending_work = None


class _GradReducer:
    """Engine shared by :class:`NaiveDDP` and :class:`MoEDP`."""

    def __init__(self, named_params: "OrderedDict[str, torch.nn.Parameter]", *, sync: bool,
                 bucket_cap_mb: float, gradient_as_bucket_view: bool, process_group,
                 reduce_op: str, num_grad_acc_iter: int, verbose: bool, group_fn=None,
                 use_symm: Optional[bool] = None):
        self.params = named_params
        self.sync = bool(sync)
        self.bucket_cap_bytes = int(bucket_cap_mb * 1024 * 1024)
        self.as_view = bool(gradient_as_bucket_view)
        self.default_group = process_group
        self.group_fn = group_fn
        op = str(reduce_op).lower()
        if op not in ("avg", "sum"):
            raise ValueError(f"reduce_op must be 'avg' or 'sum', got {reduce_op!r}")
        self.average = op == "avg"
        self.num_grad_acc_iter = max(int(num_grad_acc_iter), 1)
        self.verbose = verbose
        self.reduce_time = 0.0
        self._acc_counter: Dict[int, int] = {}
        self._hooks = []
        self._finalized = True

        any_param = next(iter(named_params.values()), None)
        self.device = any_param.device if any_param is not None else torch.device("cpu")
        self.on_cuda = self.device.type == "cuda"
        self.comm_stream = None
        if self.on_cuda:
            lo, hi = torch.cuda.Stream.priority_range()
            self.comm_stream = torch.cuda.Stream(device=self.device, priority=hi)
        self._use_symm_req = use_symm
        self.buckets: List[GradBucket] = []
        self.param_bucket: Dict[str, GradBucket] = {}
        self._claim_params()
        self._build_buckets()
        self._register_hooks()

    def _claim_params(self) -> None:
        """One owner for ``p.grad``: mark the parameters as reduced by this engine so a
        ``Bf16ZeroOptimizer`` built later turns into a consumer of the reduced gradients, and tell
        an optimizer built *earlier* to hand ownership over (hybrid ZeRO, ddp/zero_optim.py)."""
        import weakref
        me = weakref.ref(self)
        for p in self.params.values():
            if not p.requires_grad:
                continue
            zref = getattr(p, "_tdp_zero", None)
            z = zref() if zref is not None else None
            if z is not None:
                z.attach_external_reducer(self)
            p._tdp_reducer = me

    # ------------------------------------------------------------------ groups
    def _group_of(self, name: str, p: torch.nn.Parameter):
        if self.group_fn is not None:
            g = self.group_fn(name, p)
            if g is not None:
                return g
        return self.default_group

    @staticmethod
    def _group_size(group) -> int:
        return dist.get_world_size(group) if dist.is_initialized() else 1

    # ------------------------------------------------------------------ bucket construction
    def _build_buckets(self) -> None:
        """Reverse registration order approximates backward order, so the first buckets to fill
        are the first whose reduction can overlap the rest of backward."""
        trainable = [(n, p) for n, p in self.params.items() if p.requires_grad]
        plan: List[dict] = []   # {key, dtype, group, items:[(name,p)], numel}
        open_by_key: Dict[Tuple, dict] = {}
        for name, p in reversed(trainable):
            group = self._group_of(name, p)
            key = (p.dtype, p.device, id(group))
            esize = p.element_size()
            cap = max(self.bucket_cap_bytes // esize, 1)
            align = max(1, _ALIGN_BYTES // esize)
            big = p.numel() * esize >= 0.8 * self.bucket_cap_bytes
            if big:
                plan.append(dict(dtype=p.dtype, device=p.device, group=group, items=[(name, p)],
                                 numel=align_up(p.numel(), align)))
                continue
            cur = open_by_key.get(key)
            if cur is None or align_up(cur["numel"], align) + p.numel() > cap:
                cur = dict(dtype=p.dtype, device=p.device, group=group, items=[], numel=0)
                open_by_key[key] = cur
                plan.append(cur)
            cur["numel"] = align_up(cur["numel"], align) + p.numel()
            cur["items"].append((name, p))

        # symmetric memory: one allocation per (group) holding all of that group's buckets
        symm_bufs: Dict[int, Tuple[object, int]] = {}
        if self.on_cuda and self._use_symm_req is not False and dist.is_initialized():
            need: Dict[int, int] = {}
            groups: Dict[int, object] = {}
            for b in plan:
                if b["dtype"] in (torch.bfloat16, torch.float32) and self._group_size(b["group"]) > 1:
                    gid = id(b["group"])
                    need[gid] = need.get(gid, 0) + align_up(
                        align_up(b["numel"], 8) * torch.empty((), dtype=b["dtype"]).element_size(),
                        4096)
                    groups[gid] = b["group"]
            for gid, nbytes in need.items():
                sg = get_symm_group(groups[gid])
                if sg.enabled:
                    symm_bufs[gid] = [sg.alloc(nbytes), 0]
                elif self._use_symm_req:
                    raise RuntimeError(f"symmetric memory requested but unavailable: {sg.reason}")

        for i, b in enumerate(plan):
            esize = torch.empty((), dtype=b["dtype"]).element_size()
            numel = align_up(b["numel"], 8)
            buffer, symm = None, None
            entry = symm_bufs.get(id(b["group"]))
            if entry is not None and b["dtype"] in (torch.bfloat16, torch.float32) \
                    and self._group_size(b["group"]) > 1:
                sbuf, off = entry
                buffer = sbuf.view(off, (numel,), b["dtype"])
                symm = (sbuf, off)
                entry[1] = off + align_up(numel * esize, 4096)
            bucket = GradBucket(i, b["dtype"], b["device"], b["group"], numel, buffer, symm)
            for name, p in b["items"]:
                view = bucket.push(name, p.shape, p.numel())
                self.param_bucket[name] = bucket
                if self.as_view:
                    if p.grad is not None:
                        view.copy_(p.grad)
                    p.grad = view
                    if self.on_cuda and p.dtype == torch.bfloat16 and p.numel() % 8 == 0 \
                            and view.data_ptr() % 16 == 0:
                        # lets ops.linear.wgrad / colsum_param / the LayerNorm backward write the
                        # gradient straight into the bucket (weights, biases, LN parameters)
                        p._tdp_main_grad = view
                        p._tdp_grad_fresh = True
            self.buckets.append(bucket)

    # ------------------------------------------------------------------ hooks
    def _register_hooks(self) -> None:
        for name, p in self.params.items():
            if not p.requires_grad:
                continue
            self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(name)))

    def _make_hook(self, name: str):
        def hook(p: torch.Tensor):
            self._on_grad_ready(name, p)
        return hook

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for p in self.params.values():
            for attr in ("_tdp_main_grad", "_tdp_grad_fresh", "_tdp_reducer", "_tdp_grad_stream"):
                if hasattr(p, attr):
                    delattr(p, attr)

    def _on_grad_ready(self, name: str, p: torch.Tensor) -> None:
        bucket = self.param_bucket[name]
        self._finalized = False
        gs = getattr(p, "_tdp_grad_stream", None)
        if gs is not None:
            bucket.producer_streams.add(gs)
            p._tdp_grad_stream = None
        if self.as_view:
            view = bucket.views[name]
            if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                # the user dropped the view (zero_grad(set_to_none=True)): re-attach it
                view.copy_(p.grad)
                p.grad = view
        if not bucket.grad_ready():
            return
        # all grads of this bucket are final for this micro-step
        bucket._ready = 0
        cnt = self._acc_counter.get(bucket.index, 0) + 1
        if cnt < self.num_grad_acc_iter:
            self._acc_counter[bucket.index] = cnt   # still accumulating (e.g. PP micro-batches)
            return
        self._acc_counter[bucket.index] = 0
        if not self.sync:
            self._reduce_bucket(bucket)

    # ------------------------------------------------------------------ reduction
    def _pack(self, bucket: GradBucket) -> None:
        if self.as_view:
            return
        for name in bucket.names:
            g = self.params[name].grad
            if g is None:
                bucket.views[name].zero_()
            else:
                bucket.views[name].copy_(g)

    def _unpack(self, bucket: GradBucket) -> None:
        if self.as_view:
            return
        for name in bucket.names:
            p = self.params[name]
            if p.grad is not None:
                p.grad.copy_(bucket.views[name])

    def _reduce_bucket(self, bucket: GradBucket) -> None:
        if bucket.reduced:
            return
        bucket.reduced = True
        world = self._group_size(bucket.group)
        if world <= 1:
            return
        t0 = time.perf_counter() if self.verbose else 0.0
        if self.on_cuda:
            cur = torch.cuda.current_stream(self.device)
            self.comm_stream.wait_stream(cur)       # device-side dependency only
            for ps in bucket.producer_streams:      # gradients written outside autograd's view
                if ps != cur:
                    self.comm_stream.wait_stream(ps)
            bucket.producer_streams.clear()
            self._comm_used = True
            with torch.cuda.stream(self.comm_stream):
                self._pack(bucket)
                payload = bucket.payload()
                if bucket.symm is not None and bucket.fused_step is not None:
                    bucket.fused_step(bucket)
                elif bucket.symm is not None:
                    sbuf, off = bucket.symm
                    sbuf.all_reduce_(off, payload.numel(), bucket.dtype,
                                     (1.0 / world) if self.average else 1.0)
                else:
                    op = dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM
                    dist.all_reduce(payload, op=op, group=bucket.group)
                self._unpack(bucket)
        else:
            self._pack(bucket)
            payload = bucket.payload()
            bucket.pending_work = dist.all_reduce(payload, op=dist.ReduceOp.SUM,
                                                  group=bucket.group, async_op=True)
        if self.verbose:
            self.reduce_time += time.perf_counter() - t0

    def finalize(self) -> None:
        """Make every bucket's reduction visible to the compute stream; reset counters."""
        t0 = time.perf_counter() if self.verbose else 0.0
        for bucket in self.buckets:
            if not bucket.reduced:
                # sync mode, or grads that never fired (unused params): reduce now
                if self.as_view:
                    pass
                self._reduce_bucket(bucket)
        if self.on_cuda:
            # (only when something was enqueued: waiting on an idle, non-capturing stream would be
            # an illegal cross-capture dependency under CUDA-graph capture)
            if getattr(self, "_comm_used", False):
                torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
                self._comm_used = False
        else:
            for bucket in self.buckets:
                if bucket.pending_work is not None:
                    bucket.pending_work.wait()
                    if self.average:
                        bucket.payload().div_(self._group_size(bucket.group))
                    self._unpack(bucket)
        for bucket in self.buckets:
            bucket.grad_reset()
        self._acc_counter.clear()
        self._finalized = True
        for p in self.params.values():      # next backward overwrites instead of accumulating
            if hasattr(p, "_tdp_grad_fresh"):
                p._tdp_grad_fresh = True
        if self.verbose:
            self.reduce_time += time.perf_counter() - t0
            print(f"[NaiveDDP] rank {dist.get_rank() if dist.is_initialized() else 0}: "
                  f"host time in reduce this iter {self.reduce_time * 1e3:.3f} ms", flush=True)
            self.reduce_time = 0.0

    # ------------------------------------------------------------------ param broadcast
    @staticmethod
    def broadcast_tensors(tensors: Iterable[torch.Tensor], src_global_rank: int, group) -> None:
        """Broadcast many tensors with few collectives: coalesce per dtype into flat buffers
        (the reference issues one NCCL broadcast per tensor, naive_ddp.py:226-230)."""
        if not dist.is_initialized() or dist.get_world_size(group) <= 1:
            return
        by_dtype: Dict[Tuple, List[torch.Tensor]] = {}
        for t in tensors:
            by_dtype.setdefault((t.dtype, t.device), []).append(t)
        for (_, _), ts in by_dtype.items():
            chunk: List[torch.Tensor] = []
            size = 0
            def flush():
                nonlocal chunk, size
                if not chunk:
                    return
                flat = torch.cat([t.detach().reshape(-1) for t in chunk])
                dist.broadcast(flat, src_global_rank, group=group)
                off = 0
                with torch.no_grad():
                    for t in chunk:
                        t.copy_(flat[off:off + t.numel()].view_as(t))
                        off += t.numel()
                chunk, size = [], 0
            for t in ts:
                chunk.append(t)
                size += t.numel() * t.element_size()
                if size >= 256 * 1024 * 1024:
                    flush()
            flush()


class NaiveDDP(torch.nn.Module):
    """Drop-in for the reference ``NaiveDDP`` (see module docstring).

    Typical loop::

        model = NaiveDDP(model, sync=False, gradient_as_bucket_view=True)
        loss = model(x).sum(); loss.backward()
        model.reduce_gradients()
        optimizer.step(); optimizer.zero_grad()
    """

    def __init__(self, module: torch.nn.Module, sync: bool = False, bucket_cap_mb: float = 25,
                 gradient_as_bucket_view: bool = False, process_group=None, dp_rank0: int = 0,
                 reduce_op: str = "avg", **kwargs):
        super().__init__()
        self.module = module
        self.group = process_group
        self.dp_rank0 = int(dp_rank0)
        self.verbose = bool(kwargs.pop("verbose", False))
        num_grad_acc_iter = int(kwargs.pop("num_grad_acc_iter", 1))
        use_symm = kwargs.pop("use_symm", None)
        broadcast = kwargs.pop("broadcast_params", True)
        if kwargs:
            raise TypeError(f"unexpected arguments: {sorted(kwargs)}")

        ignore = set(getattr(module, "_ddp_params_and_buffers_to_ignore", []) or [])
        self.parameters_to_ignore = ignore
        named = OrderedDict((n, p) for n, p in module.named_parameters() if n not in ignore)
        if broadcast:
            self.broadcast_params()
        self.reducer = _GradReducer(named, sync=sync, bucket_cap_mb=bucket_cap_mb,
                                    gradient_as_bucket_view=gradient_as_bucket_view,
                                    process_group=process_group, reduce_op=reduce_op,
                                    num_grad_acc_iter=num_grad_acc_iter, verbose=self.verbose,
                                    group_fn=self._get_group, use_symm=use_symm)

    # overridable: per-parameter process group (reference: naive_ddp.py:95-96)
    def _get_group(self, name: str, param: torch.nn.Parameter):
        return self.group

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    @property
    def buckets(self) -> List[GradBucket]:
        return self.reducer.buckets

    def reduce_gradients(self) -> None:
        """Call after backward (of the last micro-batch): finishes / waits for all bucket
        reductions.  Afterwards every ``p.grad`` holds the group-reduced gradient."""
        self.reducer.finalize()

    def broadcast_params(self) -> None:
        """Make every replica start from ``dp_rank0``'s parameters and buffers."""
        tensors = [t for n, t in self.module.state_dict().items()
                   if n not in self.parameters_to_ignore and torch.is_tensor(t)]
        src = self.dp_rank0
        if self.group is not None and dist.is_initialized():
            # ``dp_rank0`` is a *global* rank (reference semantics, naive_ddp.py:226-230, which
            # fails when global rank 0 is not a member); fall back to the group's first rank
            ranks = dist.get_process_group_ranks(self.group)
            if src not in ranks:
                src = ranks[0]
        _GradReducer.broadcast_tensors(tensors, src, self.group)

    def sync_comm(self) -> None:
        if self.reducer.on_cuda:
            self.reducer.comm_stream.synchronize()

    def reduce_dispatch(self, name: str, p: torch.nn.Parameter) -> None:
        """What the gradient hook of parameter ``name`` runs (reference: naive_ddp.py:129-171):
        marks the gradient final for this micro-batch and launches the bucket's reduction once
        all of its gradients are.  Exposed for custom schedules that produce a gradient outside
        autograd and for subclasses; the hooks call the reducer directly."""
        self.reducer._on_grad_ready(name, p)

    def set_num_grad_acc_iter(self, n: int) -> None:
        self.reducer.num_grad_acc_iter = max(int(n), 1)

    def remove_hooks(self) -> None:
        """Detach the gradient hooks (and the direct weight-gradient attributes) from the
        parameters.  Not done automatically when the wrapper is garbage collected: the hooks keep
        the reducer alive, so ``NaiveDDP(model)`` keeps synchronising even if the caller drops the
        wrapper object and trains through ``model`` (as with the reference)."""
        self.reducer.remove_hooks()

    def zero_grad(self, set_to_none: bool = False) -> None:  # keep bucket views alive by default
        for p in self.module.parameters():
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    # (bucket views cannot be detach_()ed in place -- same rule as torch's
                    # Optimizer.zero_grad)
                    if p.grad.grad_fn is not None:
                        p.grad.detach_()
                    else:
                        p.grad.requires_grad_(False)
                    p.grad.zero_()


NaiveDdp = NaiveDDP  # spelling used by the north-star document


class MoEDP:
    """Replicated-expert data parallelism: all-reduce the gradients of *expert* parameters over
    the ``moe_dp`` group (ranks holding replicas of the same experts), while the dense parameters
    go through the ordinary :class:`NaiveDDP` with the expert names listed in
    ``module._ddp_params_and_buffers_to_ignore`` (reference: ddp/naive_ddp.py:233-441,
    ddp/moe_dp.md).  The reference's default path performs no reduction at all (all_reduce
    commented out, :316-322); this one does."""

    def __init__(self, moe_params: Dict[str, torch.nn.Parameter], moe_dp_group, moe_dp_rank0: int = 0,
                 overlap_comm: bool = True, reduce_op: str = "avg", sync: bool = False,
                 num_grad_acc_iter: int = 1, bucket_cap_mb: float = 25, verbose: bool = False,
                 use_symm: Optional[bool] = None):
        self.params = OrderedDict(moe_params)
        self.group = moe_dp_group
        self.rank0 = int(moe_dp_rank0)
        _GradReducer.broadcast_tensors([p.data for p in self.params.values()], self.rank0,
                                       self.group)
        self.reducer = _GradReducer(self.params, sync=(sync or not overlap_comm),
                                    bucket_cap_mb=bucket_cap_mb, gradient_as_bucket_view=True,
                                    process_group=moe_dp_group, reduce_op=reduce_op,
                                    num_grad_acc_iter=num_grad_acc_iter, verbose=verbose,
                                    use_symm=use_symm)

    def reduce_gradients(self) -> None:
        self.reducer.finalize()

    def broadcast_params(self) -> None:
        """Re-send the expert parameters from ``moe_dp_rank0`` (done once by the constructor)."""
        _GradReducer.broadcast_tensors([p.data for p in self.params.values()], self.rank0,
                                       self.group)

    def reduce_dispatch(self, name: str, p: torch.nn.Parameter) -> None:
        self.reducer._on_grad_ready(name, p)

    def remove_hooks(self) -> None:
        self.reducer.remove_hooks()


moe_dp_mod: Optional[MoEDP] = None


def create_moe_dp_hooks(params: Dict[str, torch.nn.Parameter], moe_dp_group, moe_dp_rank0: int = 0,
                        overlap_comm: bool = True, reduce_op: str = "avg", sync: bool = False,
                        num_grad_acc_iter: int = 1, **kwargs) -> MoEDP:
    """Install gradient hooks on the expert parameters (``{name: param}``)."""
    global moe_dp_mod
    moe_dp_mod = MoEDP(params, moe_dp_group, moe_dp_rank0, overlap_comm=overlap_comm,
                       reduce_op=reduce_op, sync=sync, num_grad_acc_iter=num_grad_acc_iter, **kwargs)
    return moe_dp_mod


def moe_dp_iter_step() -> None:
    """Call once per iteration after backward: waits for the expert-gradient reductions."""
    if moe_dp_mod is not None:
        moe_dp_mod.reduce_gradients()
